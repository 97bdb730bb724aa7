// kernels.cuh -- sm_100a device code of the IVF-RaBitQ query hot path.
//
// Written from scratch for B200; the reference (kemingy/rabitq v0.2.2, Rust + AVX2) is cited per kernel as
// file:line relative to /root/reference so parity can be checked.  Floating-point work that feeds a comparison
// or a rounding reproduces the reference's AVX evaluation order exactly (8 partial sums, lane v = elements
// v, v+8, ... with one fused multiply-add each, then ((s0+s4)+(s1+s5))+((s2+s6)+(s3+s7))); everything else is
// integer.  Compiled with -fmad=false: an FMA exists only where `fmaf` is written.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rq {

#define RQ_DEV __device__ __forceinline__
constexpr unsigned FULL = 0xffffffffu;

// f32::total_cmp / Ord32 (src/ord32.rs:12-26) as a monotone UNSIGNED key.
RQ_DEV uint32_t okey(float x) {
    uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
RQ_DEV float okey_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

// _mm256_cvtps_epi32 (src/simd.rs:215): round to nearest even; NaN and out-of-range give 0x80000000.
RQ_DEV int cvtps_epi32(float x) {
    return (fabsf(x) < 2147483648.0f) ? __float2int_rn(x) : (int)0x80000000;
}

// reduce_f32_256 (src/simd.rs:52-63, 292-303) for 8 accumulators held by one thread.
RQ_DEV float reduce8(const float* s) {
    float c0 = __fadd_rn(s[0], s[4]), c1 = __fadd_rn(s[1], s[5]);
    float c2 = __fadd_rn(s[2], s[6]), c3 = __fadd_rn(s[3], s[7]);
    return __fadd_rn(__fadd_rn(c0, c1), __fadd_rn(c2, c3));
}

// Packed fp32x2 arithmetic (Blackwell FADD2 / FFMA2): two IEEE round-to-nearest fp32 operations per instruction, each
// lane rounded exactly like the scalar op, so the AVX evaluation order is kept while the issue slots are halved.
typedef unsigned long long f32x2;
RQ_DEV f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
RQ_DEV void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
RQ_DEV f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
RQ_DEV f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
RQ_DEV float reduce8p(const f32x2* s) {  // reduce_f32_256 over 8 lanes held as 4 packed pairs (0,1) (2,3) (4,5) (6,7)
    float v[8];
    unpack2(s[0], v[0], v[1]); unpack2(s[1], v[2], v[3]); unpack2(s[2], v[4], v[5]); unpack2(s[3], v[6], v[7]);
    return reduce8(v);
}

// ---------------------------------------------------------------------------------------------------------
// K0: zero-pad queries to D (src/rabitq.rs:277-280).
__global__ void pad_queries_kernel(const float* __restrict__ q, float* __restrict__ qpad, size_t nq, int len, int D) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= nq * (size_t)D) return;
    size_t r = i / D;
    int c = (int)(i % D);
    qpad[i] = c < len ? q[r * len + c] : 0.0f;
}

// ---------------------------------------------------------------------------------------------------------
// K1: batched rotation y = q * P with the summation order of project -> simd::vector_dot_product
// (src/utils.rs:237-258, src/simd.rs:257-314): AVX lane l accumulates fma(q[r], P[r][c], acc_l) over r = l, l+8, ... in
// ascending order and reduce_f32_256 adds the 8 lanes as ((0+4)+(1+5))+((2+6)+(3+7)).  CUDA cores by necessity: the
// quantised codes downstream must be bit-exact, so y must be, so that association is reproduced (no 3xTF32).
//
// The kernel is bound by shared-memory operand bandwidth, so the register tile is made as large as the 8-lane structure
// allows: a thread owns TWO of the eight AVX lanes (one packed fma.rn.f32x2 pair) of an 8-query x 8-column tile
// (64 independent FFMA2 chains), four threads complete an output and combine their lane pairs with shuffles in the
// reference's order at the end.  Per 8 rows a thread loads 8 + 8 64-bit operands for 64 FFMA2 (2 B per FFMA2; the
// previous one-column-x-8-queries tile needed 9).  PT is P transposed (PT[c][r] = P[r][c], built once per index), so that a
// thread's two rows of a column are adjacent; tiles are staged by a two-stage cp.async pipeline (the loads of rows
// r0+64.. fly while rows r0.. are consumed); the 72-float pitch puts the 8 column groups of a warp on disjoint banks.
constexpr int ROT_TC = 64;        // output columns per CTA
constexpr int ROT_TQ = 32;        // queries per CTA (8 per warp)
constexpr int ROT_THREADS = 128;
constexpr int ROT_RCH = 64;       // rows of P per shared-memory stage
constexpr int ROT_PITCH = ROT_RCH + 8;

__global__ void transpose_square_kernel(const float* __restrict__ in, float* __restrict__ out, int D) {
    __shared__ float t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) t[i][threadIdx.x] = in[(size_t)(by + i) * D + bx + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) out[(size_t)(bx + i) * D + by + threadIdx.x] = t[threadIdx.x][i];
}

RQ_DEV void cp_async16(void* dst_smem, const void* src_gmem, bool valid) {  // 16 B global -> shared, zero-filled when !valid
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src_gmem), "r"(n) : "memory");
}
RQ_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
RQ_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int ROT_STAGE_FLOATS = (ROT_TC + ROT_TQ) * ROT_PITCH;
constexpr int ROT_SMEM_BYTES = 2 * ROT_STAGE_FLOATS * 4;  // two stages (cp.async double buffer): 55 KB

__global__ void __launch_bounds__(ROT_THREADS, 3) rotate_kernel(const float* __restrict__ qpad, const float* __restrict__ PT,
                                                                float* __restrict__ y, int nq, int D) {
    extern __shared__ __align__(16) float rot_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v = lane & 3, g = lane >> 2;  // v: AVX lanes 2v, 2v+1;  g: columns g, g+8, ..., g+56 of the CTA tile
    const int col0 = blockIdx.x * ROT_TC, qb = blockIdx.y * ROT_TQ;
    auto load_stage = [&](int stage, int r0) {  // P tile (64 columns x 64 rows) and query tile (32 x 64) of rows [r0, r0 + 64)
        float* sP = rot_smem + stage * ROT_STAGE_FLOATS;
        float* sq = sP + ROT_TC * ROT_PITCH;
#pragma unroll
        for (int i = 0; i < (ROT_TC * ROT_RCH / 4) / ROT_THREADS; i++) {
            const int idx = tid + i * ROT_THREADS, c = idx / (ROT_RCH / 4), r4 = idx % (ROT_RCH / 4);
            cp_async16(sP + c * ROT_PITCH + r4 * 4, PT + (size_t)(col0 + c) * D + r0 + r4 * 4, true);
        }
#pragma unroll
        for (int i = 0; i < (ROT_TQ * ROT_RCH / 4) / ROT_THREADS; i++) {
            const int idx = tid + i * ROT_THREADS, row = idx / (ROT_RCH / 4), r4 = idx % (ROT_RCH / 4);
            const bool ok = qb + row < nq;
            cp_async16(sq + row * ROT_PITCH + r4 * 4, qpad + (size_t)(ok ? qb + row : 0) * D + r0 + r4 * 4, ok);
        }
        cp_async_commit();
    };
    f32x2 acc[8][8];
#pragma unroll
    for (int t = 0; t < 8; t++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[t][j] = 0ull;
    const int nch = D / ROT_RCH;
    load_stage(0, 0);
    for (int it = 0; it < nch; it++) {
        if (it + 1 < nch) {
            load_stage((it + 1) & 1, (it + 1) * ROT_RCH);  // in flight while this stage is consumed
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* sP = rot_smem + (it & 1) * ROT_STAGE_FLOATS + (size_t)g * ROT_PITCH + 2 * v;
        const float* sq = rot_smem + (it & 1) * ROT_STAGE_FLOATS + ROT_TC * ROT_PITCH + (size_t)(warp * 8) * ROT_PITCH + 2 * v;
#pragma unroll 2
        for (int r = 0; r < ROT_RCH; r += 8) {
            f32x2 pv[8], qv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) pv[j] = *reinterpret_cast<const f32x2*>(sP + 8 * j * ROT_PITCH + r);
#pragma unroll
            for (int t = 0; t < 8; t++) qv[t] = *reinterpret_cast<const f32x2*>(sq + t * ROT_PITCH + r);
#pragma unroll
            for (int t = 0; t < 8; t++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[t][j] = fma2(qv[t], pv[j], acc[t][j]);
        }
        __syncthreads();  // the stage is refilled by the next iteration's load
    }
    // reduce_f32_256: c_i = s_i + s_{i+4}; (c0 + c1) + (c2 + c3).  Thread v holds s_{2v}, s_{2v+1}; all four threads of a
    // group end up with the same bits (fp32 addition commutes), thread v stores columns j = 2v, 2v+1.
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int q = qb + warp * 8 + t;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float lo, hi, olo, ohi;
            unpack2(acc[t][j], lo, hi);
            unpack2(__shfl_xor_sync(FULL, acc[t][j], 2), olo, ohi);
            const float s = __fadd_rn(__fadd_rn(lo, olo), __fadd_rn(hi, ohi));  // v even: c0 + c1, v odd: c2 + c3
            const float res = __fadd_rn(s, __shfl_xor_sync(FULL, s, 1));
            if ((j >> 1) == v && q < nq) y[(size_t)q * D + col0 + g + 8 * j] = res;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2: squared L2 of every rotated centroid to every rotated query, order of simd::l2_squared_distance
// (src/rabitq.rs:283-293, src/simd.rs:14-73: diff = c - y rounded, then sum = fma(diff, diff, sum)).
// CTA tile = 64 centroids x 32 queries; both operand tiles are staged in shared memory with coalesced 128-bit
// loads (the centroid tile padded to a 68-float pitch so that 32 threads reading 32 different rows with LDS.128
// hit disjoint banks).  Thread = one centroid x 8 queries, 8 AVX-lane accumulators each (64 registers).
constexpr int CD_TQ = 8;        // queries per thread
constexpr int CD_QG = 4;        // query groups per CTA
constexpr int CD_TC = 64;       // centroids per CTA
constexpr int CD_THREADS = CD_TC * CD_QG;
constexpr int CD_DCH = 64;
constexpr int CD_PITCH = CD_DCH + 4;

// `run_if`: NULL = always; otherwise the kernel runs only when *run_if != 0 (the device-side fallback of the tensor-core
// prefilter, prefilter.cuh: no host round trip decides whether the classic path is needed).
__global__ void __launch_bounds__(CD_THREADS, 2) centroid_dist_kernel(const float* __restrict__ cent, const float* __restrict__ y,
                                                                      float* __restrict__ out, int nq, int K, int D,
                                                                      const uint32_t* __restrict__ run_if) {
    if (run_if && *run_if == 0u) return;
    __shared__ __align__(16) float sc[CD_TC][CD_PITCH];
    __shared__ float4 sy[CD_QG * CD_TQ][CD_DCH / 4];
    const int tid = threadIdx.x;
    const int cl = tid & (CD_TC - 1), qg = tid / CD_TC;
    // tiles (64 centroids x 32 queries) are walked with a grid-stride loop: as the prefilter's fallback the kernel is launched
    // with a small grid, so the usual case (flag clear) costs a few hundred exiting CTAs instead of one per tile
    const int tiles_c = (K + CD_TC - 1) / CD_TC, tiles_q = (nq + CD_QG * CD_TQ - 1) / (CD_QG * CD_TQ);
    for (long long tile = blockIdx.x; tile < (long long)tiles_c * tiles_q; tile += gridDim.x) {
    const int c0 = (int)(tile % tiles_c) * CD_TC, c = c0 + cl;
    const int qb = (int)(tile / tiles_c) * (CD_QG * CD_TQ), q0 = qb + qg * CD_TQ;
    f32x2 acc[CD_TQ][4];
#pragma unroll
    for (int t = 0; t < CD_TQ; t++)
#pragma unroll
        for (int v = 0; v < 4; v++) acc[t][v] = 0ull;
    for (int d0 = 0; d0 < D; d0 += CD_DCH) {
#pragma unroll
        for (int i = 0; i < (CD_TC * CD_DCH / 4) / CD_THREADS; i++) {
            const int idx = tid + i * CD_THREADS, row = idx / (CD_DCH / 4), c4 = idx % (CD_DCH / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + row < K) v = __ldg(reinterpret_cast<const float4*>(cent + (size_t)(c0 + row) * D + d0) + c4);
            *reinterpret_cast<float4*>(&sc[row][c4 * 4]) = v;
        }
#pragma unroll
        for (int i = 0; i < (CD_QG * CD_TQ * CD_DCH / 4) / CD_THREADS; i++) {
            const int idx = tid + i * CD_THREADS, row = idx / (CD_DCH / 4), c4 = idx % (CD_DCH / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (qb + row < nq) v = __ldg(reinterpret_cast<const float4*>(y + (size_t)(qb + row) * D + d0) + c4);
            sy[row][c4] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int d = 0; d < CD_DCH; d += 8) {
            const ulonglong2 ca = *reinterpret_cast<const ulonglong2*>(&sc[cl][d]), cb = *reinterpret_cast<const ulonglong2*>(&sc[cl][d + 4]);
#pragma unroll
            for (int t = 0; t < CD_TQ; t++) {
                const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(&sy[qg * CD_TQ + t][d / 4]);
                const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(&sy[qg * CD_TQ + t][d / 4 + 1]);
                f32x2 f;
                f = sub2(ca.x, a.x); acc[t][0] = fma2(f, f, acc[t][0]);  // AVX lanes 0,1: diff = c - y, then fma(diff, diff, sum)
                f = sub2(ca.y, a.y); acc[t][1] = fma2(f, f, acc[t][1]);  // 2,3
                f = sub2(cb.x, b.x); acc[t][2] = fma2(f, f, acc[t][2]);  // 4,5
                f = sub2(cb.y, b.y); acc[t][3] = fma2(f, f, acc[t][3]);  // 6,7
            }
        }
        __syncthreads();
    }
    if (c < K) {
#pragma unroll
        for (int t = 0; t < CD_TQ; t++)
            if (q0 + t < nq) out[(size_t)(q0 + t) * K + c] = reduce8p(acc[t]);
    }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Block-wide exclusive scan of n u32 values in shared memory (in place); returns the total.
template <int THREADS>
RQ_DEV uint32_t block_exclusive_scan(uint32_t* a, int n, uint32_t* warp_tot /* THREADS/32 + 1 */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + THREADS - 1) / THREADS;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    uint32_t s = 0;
    for (int i = lo; i < hi; i++) s += a[i];
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < THREADS / 32 ? warp_tot[lane] : 0, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(FULL, winc, o);
            if (lane >= o) winc += v;
        }
        if (lane < THREADS / 32) warp_tot[lane] = winc - w;
        if (lane == THREADS / 32 - 1) warp_tot[THREADS / 32] = winc;
    }
    __syncthreads();
    uint32_t run = warp_tot[warp] + inc - s;
    for (int i = lo; i < hi; i++) {
        uint32_t v = a[i];
        a[i] = run;
        run += v;
    }
    uint32_t total = warp_tot[THREADS / 32];
    __syncthreads();
    return total;
}

// ---------------------------------------------------------------------------------------------------------
// K2b: top-`P` nearest centroids per query, ascending (src/rabitq.rs:294-297: select_nth_unstable_by +
// truncate + sort_by(total_cmp)).  Ties between EQUAL distances are unspecified in the reference; here the
// smaller centroid index wins.  Also emits, per (query, rank), the prefix of 32-vector words of the probed
// clusters (the survivor-slot layout) and the per-query totals (`rough` counter of src/rerank.rs:105).
constexpr int SEL_THREADS = 256;
constexpr int SEL_SAMPLE = 512;   // pivot path: sample size
constexpr int SEL_CAND = 1024;    // pivot path: candidate capacity (sel_buf holds max(Ppow2, SEL_CAND) entries)

__global__ void __launch_bounds__(SEL_THREADS) select_probe_kernel(const float* __restrict__ cdist, int K, int P, int Ppow2, int cache_keys,
                                                                   const uint32_t* __restrict__ offsets,
                                                                   const uint32_t* __restrict__ offsets_g,
                                                                   uint32_t* __restrict__ probe_ids, float* __restrict__ probe_dist,
                                                                   uint32_t* __restrict__ slot_local, uint32_t* __restrict__ q_words,
                                                                   uint32_t* __restrict__ q_pairs, uint32_t* __restrict__ q_p0,
                                                                   const uint32_t* __restrict__ run_if) {
    if (run_if && *run_if == 0u) return;
    extern __shared__ unsigned long long sel_buf[];  // Ppow2 (key,index) pairs, then (optionally) the K keys of this query
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_bucket, s_need, s_nout, s_p0;
    __shared__ uint32_t warp_tot[SEL_THREADS / 32 + 1];
    __shared__ uint32_t eq_cnt[SEL_THREADS];
    const int tid = threadIdx.x, lane = tid & 31;
    const size_t q = blockIdx.x;
    const float* row = cdist + q * (size_t)K;
    const int buf_n = max(Ppow2, SEL_CAND);
    uint32_t* samp = reinterpret_cast<uint32_t*>(sel_buf + buf_n);      // SEL_SAMPLE keys
    uint32_t* skeys = samp + SEL_SAMPLE;                                // K keys when cached
    // Keys are taken relative to the smallest one: squared distances share their high bits, and a radix digit that is
    // identical for every key would serialise all histogram updates on one shared-memory bin.
    if (tid == 0) s_p0 = 0xffffffffu;
    // ---- fast path: pivot from a sorted sample, one collecting pass, sort of the few candidates ----------------
    // Exact: the result is the P smallest (key, index) pairs, same as the radix path, which remains the fallback when
    // the pivot admits fewer than P or more than SEL_CAND keys.
    bool have_result = false;
    if (K >= 16 * P && K >= 4 * SEL_SAMPLE) {
        const int stride = K / SEL_SAMPLE;
        for (int i = tid; i < SEL_SAMPLE; i += SEL_THREADS) samp[i] = okey(__ldg(&row[i * stride]));
        if (tid == 0) s_nout = 0;
        __syncthreads();
        for (int k2 = 2; k2 <= SEL_SAMPLE; k2 <<= 1)
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < SEL_SAMPLE; i += SEL_THREADS) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint32_t x = samp[i], z = samp[ixj];
                        const bool up = (i & k2) == 0;
                        if ((x > z) == up) { samp[i] = z; samp[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        const int t = min(SEL_SAMPLE - 1, 3 * ((P * SEL_SAMPLE + K - 1) / K) + 8);
        const uint32_t pivot = samp[t];
        auto take = [&](float v, int i) {
            const uint32_t key = okey(v);
            if (key <= pivot) {
                const uint32_t pos = atomicAdd(&s_nout, 1u);
                if (pos < (uint32_t)SEL_CAND) sel_buf[pos] = ((unsigned long long)key << 32) | (uint32_t)i;
            }
        };
        if ((K & 3) == 0) {  // one pass over the keys, 128-bit loads, several in flight per thread
            const float4* row4 = reinterpret_cast<const float4*>(row);
#pragma unroll 4
            for (int i4 = tid; i4 < K / 4; i4 += SEL_THREADS) {
                const float4 v = __ldg(&row4[i4]);
                take(v.x, 4 * i4); take(v.y, 4 * i4 + 1); take(v.z, 4 * i4 + 2); take(v.w, 4 * i4 + 3);
            }
        } else {
            for (int i = tid; i < K; i += SEL_THREADS) take(__ldg(&row[i]), i);
        }
        __syncthreads();
        const int cn = (int)s_nout;
        if (cn >= P && cn <= SEL_CAND) {
            int cpow2 = 1;
            while (cpow2 < cn) cpow2 <<= 1;
            for (int i = cn + tid; i < cpow2; i += SEL_THREADS) sel_buf[i] = ~0ull;
            __syncthreads();
            for (int k2 = 2; k2 <= cpow2; k2 <<= 1)
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < cpow2; i += SEL_THREADS) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const unsigned long long x = sel_buf[i], z = sel_buf[ixj];
                            const bool up = (i & k2) == 0;
                            if ((x > z) == up) { sel_buf[i] = z; sel_buf[ixj] = x; }
                        }
                    }
                    __syncthreads();
                }
            have_result = true;
        }
        __syncthreads();
    }
    if (!have_result) {
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int i = tid; i < K; i += SEL_THREADS) {
        const uint32_t key = okey(row[i]);
        if (cache_keys) skeys[i] = key;
        kmin = min(kmin, key);
        kmax = max(kmax, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(FULL, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(FULL, kmax, o));
    }
    if (lane == 0) { hist[tid >> 5] = kmin; hist[8 + (tid >> 5)] = kmax; }
    __syncthreads();
    kmin = hist[0]; kmax = hist[8];
#pragma unroll
    for (int w = 1; w < SEL_THREADS / 32; w++) { kmin = min(kmin, hist[w]); kmax = max(kmax, hist[8 + w]); }
    __syncthreads();
    auto getkey = [&](int i) -> uint32_t { return (cache_keys ? skeys[i] : okey(row[i])) - kmin; };
    const uint32_t range = kmax - kmin;
    const int passes = range ? (32 - __clz(range) + 7) / 8 : 0;

    uint32_t prefix = 0, mask = 0, need = (uint32_t)P;
    if (tid == 0) s_nout = 0;
    for (int pass = passes - 1; pass >= 0; pass--) {
        hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < K; i += SEL_THREADS) {
            const uint32_t key = getkey(i);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t cum = 0;
            int b = 0;
            for (; b < 256; b++) {
                if (cum + hist[b] >= need) break;
                cum += hist[b];
            }
            s_bucket = (uint32_t)b;
            s_need = need - cum;
        }
        __syncthreads();
        prefix |= s_bucket << (8 * pass);
        mask |= 0xffu << (8 * pass);
        need = s_need;
        __syncthreads();
    }
    __syncthreads();
    const uint32_t kth = prefix, need_eq = need, n_less = (uint32_t)P - need_eq;
    // strictly smaller keys: any order (sorted below)
    for (int i = tid; i < K; i += SEL_THREADS) {
        uint32_t key = getkey(i);
        if (key < kth) {
            uint32_t pos = atomicAdd(&s_nout, 1u);
            sel_buf[pos] = ((unsigned long long)key << 32) | (uint32_t)i;
        }
    }
    // keys equal to the P-th: the first `need_eq` by index
    {
        const int per = (K + SEL_THREADS - 1) / SEL_THREADS;
        const int lo = min(K, tid * per), hi = min(K, lo + per);
        uint32_t c = 0;
        for (int i = lo; i < hi; i++) c += (getkey(i) == kth);
        eq_cnt[tid] = c;
        __syncthreads();
        block_exclusive_scan<SEL_THREADS>(eq_cnt, SEL_THREADS, warp_tot);
        uint32_t base = eq_cnt[tid];
        for (int i = lo; i < hi && base < need_eq; i++)
            if (getkey(i) == kth) {
                sel_buf[n_less + base] = ((unsigned long long)kth << 32) | (uint32_t)i;
                base++;
            }
    }
    for (int i = P + tid; i < Ppow2; i += SEL_THREADS) sel_buf[i] = ~0ull;
    __syncthreads();
    // bitonic sort ascending on (key, index)
    for (int k2 = 2; k2 <= Ppow2; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Ppow2; i += SEL_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = sel_buf[i], b = sel_buf[ixj];
                    bool up = (i & k2) == 0;
                    if ((a > b) == up) { sel_buf[i] = b; sel_buf[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    // outputs; the words-per-slot array overlays the head of sel_buf (as u32, after the keys are consumed)
    uint32_t my_ids[16];  // P <= 4096 -> at most 16 per thread
    int cnt = 0;
    uint32_t pairs_local = 0;
    for (int p = tid; p < P; p += SEL_THREADS) my_ids[cnt++] = (uint32_t)(sel_buf[p] & 0xffffffffu);
    __syncthreads();
    uint32_t* words = reinterpret_cast<uint32_t*>(sel_buf);
    cnt = 0;
    for (int p = tid; p < P; p += SEL_THREADS) {
        uint32_t id = my_ids[cnt++];
        probe_ids[q * P + p] = id;
        probe_dist[q * P + p] = row[id];
        uint32_t n_c = offsets[id + 1] - offsets[id];
        words[p] = (n_c + 31u) >> 5;
        // offsets_g (distributed pipeline): the pair count and the first non-empty rank refer to the WHOLE index, so that
        // every shard cuts the visit order into the same rounds; otherwise they refer to this handle's rows
        const uint32_t n_g = offsets_g ? offsets_g[id + 1] - offsets_g[id] : n_c;
        pairs_local += n_g;
        if (n_g) atomicMin(&s_p0, (uint32_t)p);  // first probe rank that holds vectors
    }
    __syncthreads();
    uint32_t total_words = block_exclusive_scan<SEL_THREADS>(words, P, warp_tot);
    for (int p = tid; p < P; p += SEL_THREADS) slot_local[q * P + p] = words[p];
    // per-query pair count
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs_local += __shfl_xor_sync(FULL, pairs_local, o);
    __syncthreads();
    if ((tid & 31) == 0) warp_tot[tid >> 5] = pairs_local;
    __syncthreads();
    if (tid == 0) {
        uint32_t s = 0;
        for (int w = 0; w < SEL_THREADS / 32; w++) s += warp_tot[w];
        q_pairs[q] = s;
        q_words[q] = total_words;
        q_p0[q] = s_p0 == 0xffffffffu ? 0u : s_p0;
    }
}

// K2b for probe > 4096 (the reference has no cap: `probe.min(k)`, src/rabitq.rs:294): the shared-memory selection above holds
// at most 4096 results per query, so this path sorts ALL K (key, index) pairs of a query in a global-memory scratch row (bitonic,
// one CTA per query) and emits the first P.  Same order as the fast path: ascending distance, smaller centroid id on ties.  Only
// brute-force-like calls come here (every query scans thousands of clusters): correctness matters, speed does not.
__global__ void __launch_bounds__(SEL_THREADS) select_probe_large_kernel(const float* __restrict__ cdist, int K, int Kpow2, int P,
                                                                         const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ offsets_g,
                                                                         unsigned long long* __restrict__ scratch /* nq x Kpow2 */,
                                                                         uint32_t* __restrict__ probe_ids, float* __restrict__ probe_dist,
                                                                         uint32_t* __restrict__ slot_local, uint32_t* __restrict__ q_words,
                                                                         uint32_t* __restrict__ q_pairs, uint32_t* __restrict__ q_p0) {
    __shared__ uint32_t warp_tot[SEL_THREADS / 32 + 1];
    __shared__ uint32_t part[SEL_THREADS];
    __shared__ uint32_t s_p0, s_pairs;
    const int tid = threadIdx.x;
    const size_t q = blockIdx.x;
    const float* row = cdist + q * (size_t)K;
    unsigned long long* buf = scratch + q * (size_t)Kpow2;
    for (int i = tid; i < Kpow2; i += SEL_THREADS) buf[i] = i < K ? (((unsigned long long)okey(row[i]) << 32) | (uint32_t)i) : ~0ull;
    if (tid == 0) { s_p0 = 0xffffffffu; s_pairs = 0u; }
    __syncthreads();
    for (int k2 = 2; k2 <= Kpow2; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Kpow2; i += SEL_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = buf[i], z = buf[ixj];
                    const bool up = (i & k2) == 0;
                    if ((x > z) == up) { buf[i] = z; buf[ixj] = x; }
                }
            }
            __syncthreads();
        }
    // outputs: ids / distances, then the exclusive prefix of 32-vector words (thread t owns a contiguous run of ranks)
    const int per = (P + SEL_THREADS - 1) / SEL_THREADS, lo = min(P, tid * per), hi = min(P, lo + per);
    uint32_t words = 0, pairs = 0, first = 0xffffffffu;
    for (int p = lo; p < hi; p++) {
        const uint32_t id = (uint32_t)(buf[p] & 0xffffffffu);
        probe_ids[q * P + p] = id;
        probe_dist[q * P + p] = row[id];
        const uint32_t n_c = offsets[id + 1] - offsets[id];
        const uint32_t n_g = offsets_g ? offsets_g[id + 1] - offsets_g[id] : n_c;
        words += (n_c + 31u) >> 5;
        pairs += n_g;
        if (n_g && first == 0xffffffffu) first = (uint32_t)p;
    }
    part[tid] = words;
    atomicAdd(&s_pairs, pairs);
    atomicMin(&s_p0, first);
    __syncthreads();
    const uint32_t total_words = block_exclusive_scan<SEL_THREADS>(part, SEL_THREADS, warp_tot);
    uint32_t run = part[tid];
    for (int p = lo; p < hi; p++) {
        slot_local[q * P + p] = run;
        const uint32_t id = probe_ids[q * P + p];
        run += (offsets[id + 1] - offsets[id] + 31u) >> 5;
    }
    if (tid == 0) {
        q_pairs[q] = s_pairs;
        q_words[q] = total_words;
        q_p0[q] = s_p0 == 0xffffffffu ? 0u : s_p0;
    }
}

// Exclusive scan over the per-query word/pair totals (single block).  q_wbase[nq] = total words.
__global__ void __launch_bounds__(1024) query_base_scan_kernel(const uint32_t* __restrict__ q_words, const uint32_t* __restrict__ q_pairs,
                                                               int nq, uint32_t* __restrict__ q_wbase,
                                                               unsigned long long* __restrict__ q_pbase, uint32_t cap_words = 0u,
                                                               uint32_t* __restrict__ spec_fail = nullptr,
                                                               uint32_t* __restrict__ tot = nullptr /* [0] words, [2..3] pairs: the block the host reads */) {
    __shared__ unsigned long long wtot[33], ptot[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (nq + 1023) / 1024;
    const int lo = min(nq, tid * per), hi = min(nq, lo + per);
    unsigned long long sw = 0, sp = 0;
    for (int i = lo; i < hi; i++) { sw += q_words[i]; sp += q_pairs[i]; }
    unsigned long long iw = sw, ip = sp;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long a = __shfl_up_sync(FULL, iw, o), b = __shfl_up_sync(FULL, ip, o);
        if (lane >= o) { iw += a; ip += b; }
    }
    if (lane == 31) { wtot[warp] = iw; ptot[warp] = ip; }
    __syncthreads();
    if (warp == 0) {
        unsigned long long a = wtot[lane], b = ptot[lane], ia = a, ib = b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long x = __shfl_up_sync(FULL, ia, o), z = __shfl_up_sync(FULL, ib, o);
            if (lane >= o) { ia += x; ib += z; }
        }
        wtot[lane] = ia - a; ptot[lane] = ib - b;
        if (lane == 31) { wtot[32] = ia; ptot[32] = ib; }
    }
    __syncthreads();
    unsigned long long rw = wtot[warp] + iw - sw, rp = ptot[warp] + ip - sp;
    for (int i = lo; i < hi; i++) {
        q_wbase[i] = (uint32_t)rw;
        q_pbase[i] = rp;
        rw += q_words[i];
        rp += q_pairs[i];
    }
    if (tid == 0) {
        q_wbase[nq] = (uint32_t)wtot[32];
        q_pbase[nq] = ptot[32];
        if (tot) {
            tot[0] = (uint32_t)wtot[32];
            *reinterpret_cast<unsigned long long*>(tot + 2) = ptot[32];
        }
        // speculative slot sizing (no host round trip in the middle of the batch): the survivor slots were sized from earlier batches;
        // a batch that needs more raises the flag, its scan gets no work and its replay empty windows, and the host repeats it
        if (spec_fail && wtot[32] > (unsigned long long)cap_words) *spec_fail = 1u;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3: per (query, probed cluster): residual, min/max, 4-bit scalar quantisation
// (src/rabitq.rs:305-317 -> src/simd.rs:117-173, 185-247).  One warp per (q, p).  The reference goes on to split the
// 4-bit codes into bit-planes (vector_binarize_query, src/simd.rs:83-107) because its inner product is AND + popcount; here
// the inner product  abdp = sum_d bit_d * (q_u[d] & 15)  (the same integer, src/utils.rs:113-135) runs on the tensor cores as a
// u8 x u8 -> s32 contraction (K4), so the record keeps the codes as BYTES, already in the order and scale K4's B fragments want:
//   u8[0 .. D)        byte rec_pos(d) = (q_u[d] & 15) << (3 - (d & 3))
//   f32[D/4 + 0..3]   lower_bound, delta, (f32) sum q_u, y_c_distance_square
//   f32[D/4 + 4]      sqrt(y_c_distance_square)
//   u32[D/4 + 5]      sum q_u (raw u32)     u32[+6] first 32-vector word of this slot     u32[+7] cluster id
// Why the scale: K4 expands a 32-bit code word w into the A fragment with ONE AND per register, a = w & (0x01010101 << t)
// (t = lane & 3), so the byte of bit 8i+t holds 2^t instead of 1; the matching query byte carries 2^(3-t), every product is
// 8 * bit * q and the accumulator is 8 * abdp, exactly (integers, |acc| <= 120 * D).
constexpr float SCALAR_1_15 = 1.0f / 15.0f;  // src/consts.rs:10
constexpr int REC_META_BYTES = 32;

// byte position of dimension d inside a record: k-step j = d / 32 (one 32-bit code word = one m16n8k32 step), class t = d & 3
// (the quad lane that owns the element), half h = (d >> 2) & 1 (fragment register b0 / b1), i = (d & 31) >> 3 (byte inside the
// register).  Two k-steps are interleaved so that a lane's bytes of steps 2jj and 2jj+1 are one 128-bit load.
__host__ __device__ __forceinline__ int rec_pos(int d) {
    const int j = d >> 5, r = d & 31;
    return 64 * (j >> 1) + 16 * (r & 3) + 8 * (j & 1) + 4 * ((r >> 2) & 1) + (r >> 3);
}

// K3 walks ONE ROUND'S INVERTED LIST (cluster -> the (q, p) items probing it, bucket_fill_kernel) and writes the record of
// list slot i to qrec + i * rec_stride: the records of a cluster are then contiguous in HBM, in the order K4 consumes them,
// and K4's producer stages all the records of a work item with a single TMA bulk copy.  (A rank-0 record lives in both the
// first-chunk round's list and the main round's list and is simply computed for each.)  rec_stride is K4's shared-memory
// record pitch (scan_rec_pitch), so the HBM image is the shared-memory image.
template <int W32T>  // W32T > 0: D = 32*W32T, the centroid is held in registers across consecutive slots of one cluster; 0: any D
__global__ void __launch_bounds__(128) quantize_kernel(const float* __restrict__ y, const float* __restrict__ cent,
                                                       const uint2* __restrict__ cl_items, const uint32_t* __restrict__ n_items_ptr,
                                                       const float* __restrict__ probe_dist,
                                                       const uint32_t* __restrict__ slot_local, const uint32_t* __restrict__ q_wbase,
                                                       const uint32_t* __restrict__ offsets, const float* __restrict__ bias,
                                                       unsigned char* __restrict__ qrec, int rec_stride, int P, int D, int pch) {
    const int lane = threadIdx.x & 31;
    const size_t wid = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const size_t n_items = *n_items_ptr;
    const size_t i_begin = wid * (size_t)pch, i_end = min(n_items, i_begin + (size_t)pch);
    if (i_begin >= n_items) return;
    const int W32 = W32T > 0 ? W32T : D / 32, RS = (D + REC_META_BYTES) / 4;  // record words (without the pitch padding)
    // the record is assembled in shared memory and leaves with coalesced 128-bit stores
    extern __shared__ __align__(16) uint32_t qz_smem[];
    uint32_t* sr = qz_smem + (size_t)(threadIdx.x >> 5) * RS;
    unsigned char* sb = reinterpret_cast<unsigned char*>(sr);
    // this lane's byte of k-step g sits at sb[lane_pos + 64 * (g >> 1) + 8 * (g & 1)]   (rec_pos with r = lane)
    const int lane_pos = 16 * (lane & 3) + 4 * ((lane >> 2) & 1) + (lane >> 3);
    const int lane_shift = 3 - (lane & 3);
    float cv[W32T > 0 ? W32T : 1];
    uint32_t c_held = 0xffffffffu;
    // lane l fetches everything slot i_begin + l needs besides the two rows: one dependent chain for the warp's slots, not one per slot
    uint2 my_item = make_uint2(0u, 0u);
    float my_ycd = 0.f;
    uint32_t my_wb = 0u, my_skip = 1u;
    if (i_begin + lane < i_end) {
        my_item = __ldg(&cl_items[i_begin + lane]);  // (q * P + p, cluster)
        my_ycd = probe_dist[my_item.x];
        my_wb = q_wbase[my_item.x / (uint32_t)P] + slot_local[my_item.x];
        my_skip = (offsets && offsets[my_item.y + 1] == offsets[my_item.y]) ? 1u : 0u;  // cluster held by another shard: its record is never read
    }
    for (size_t slot = i_begin; slot < i_end; slot++) {
    const int sl = (int)(slot - i_begin);
    const size_t gw = __shfl_sync(FULL, my_item.x, sl);
    const uint32_t c = __shfl_sync(FULL, my_item.y, sl);
    const float ycd = __shfl_sync(FULL, my_ycd, sl);
    const uint32_t wb = __shfl_sync(FULL, my_wb, sl);
    if (__shfl_sync(FULL, my_skip, sl)) continue;
    const size_t q = gw / (size_t)P;
    const float* yr = y + q * (size_t)D;
    const float* cr = cent + (size_t)c * D;
    uint32_t* rec = reinterpret_cast<uint32_t*>(qrec + slot * (size_t)rec_stride);
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    float rr[W32T > 0 ? W32T : 1];
    if constexpr (W32T > 0) {
        if (c != c_held) {  // warp-uniform
#pragma unroll
            for (int g = 0; g < W32T; g++) cv[g] = __ldg(&cr[g * 32 + lane]);
            c_held = c;
        }
#pragma unroll
        for (int g = 0; g < W32T; g++) rr[g] = __fsub_rn(__ldg(&yr[g * 32 + lane]), cv[g]);
#pragma unroll
        for (int g = 0; g < W32T; g++) { mn = fminf(mn, rr[g]); mx = fmaxf(mx, rr[g]); }
    } else {
        for (int d = lane; d < D; d += 32) {
            float r = __fsub_rn(yr[d], cr[d]);
            mn = fminf(mn, r);
            mx = fmaxf(mx, r);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    }
    const float delta = __fmul_rn(__fsub_rn(mx, mn), SCALAR_1_15);  // rabitq.rs:307
    const float inv = __fdiv_rn(1.0f, delta);                       // :308 recip()
    int sum = 0;
    auto emit = [&](int g, float r) {
        int qi;
        if (bias) {  // scalar_quantize_raw (src/utils.rs:194-209): `((v - lo) * mul + bias[i]) as u8` = truncate, saturate, NaN -> 0
            const float f = __fadd_rn(__fmul_rn(__fsub_rn(r, mn), inv), __ldg(&bias[g * 32 + lane]));
            qi = (f != f) ? 0 : (f <= 0.0f ? 0 : (f >= 255.0f ? 255 : (int)f));
        } else {
            qi = cvtps_epi32(__fmul_rn(__fsub_rn(r, mn), inv));
        }
        sum += qi;  // i32 lanes wrap like _mm256_add_epi32
        // only bits 0..3 of the stored byte reach the bit-planes of the reference (src/simd.rs:95-104)
        sb[lane_pos + 64 * (g >> 1) + 8 * (g & 1)] = (unsigned char)((qi & 15) << lane_shift);
    };
    if constexpr (W32T > 0) {
#pragma unroll
        for (int g = 0; g < W32T; g++) emit(g, rr[g]);
    } else {
        for (int g = 0; g < W32; g++) emit(g, __fsub_rn(yr[g * 32 + lane], cr[g * 32 + lane]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    if (lane == 0) {
        float* rf = reinterpret_cast<float*>(sr + D / 4);
        rf[0] = mn;
        rf[1] = delta;
        rf[2] = __uint2float_rn((uint32_t)sum);  // `scalar_sum as f32`, rabitq.rs:322
        rf[3] = ycd;
        rf[4] = __fsqrt_rn(ycd);                 // rabitq.rs:346
        sr[D / 4 + 5] = (uint32_t)sum;
        sr[D / 4 + 6] = wb;
        sr[D / 4 + 7] = c;
    }
    __syncwarp();
    for (int i = lane; i < RS / 4; i += 32) reinterpret_cast<uint4*>(rec)[i] = reinterpret_cast<const uint4*>(sr)[i];
    __syncwarp();  // the staging area is rewritten by the next record
    }
}

// K3 for dim <= 256: EIGHT lanes per record, four records per warp side by side.  A lane owns the 16-byte chunks l8, l8 + 8, ... of
// the rotated query row (dims 32 i + 4 l8 + e), so a group's loads are 128 contiguous bytes and its min / max / sum reductions are
// three shuffle steps instead of five; the centroid row stays in registers across the consecutive slots of one cluster.  Same
// arithmetic per element as quantize_kernel (the reductions are order-free: min, max, integer sum), same record.
template <int W32T>
__global__ void __launch_bounds__(128) quantize_small_kernel(const float* __restrict__ y, const float* __restrict__ cent,
                                                             const uint2* __restrict__ cl_items, const uint32_t* __restrict__ n_items_ptr,
                                                             const float* __restrict__ probe_dist, const uint32_t* __restrict__ slot_local,
                                                             const uint32_t* __restrict__ q_wbase, const uint32_t* __restrict__ offsets,
                                                             const float* __restrict__ bias, unsigned char* __restrict__ qrec, int rec_stride,
                                                             int P, int pch) {
    constexpr int D = 32 * W32T, RS = (D + REC_META_BYTES) / 4;
    const int lane = threadIdx.x & 31, l8 = lane & 7, grp = lane >> 3;
    const size_t gid = ((size_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * 4 + grp;  // group index
    const size_t n_items = *n_items_ptr;
    const size_t i_begin = min(n_items, gid * (size_t)pch), i_end = min(n_items, i_begin + (size_t)pch);
    // (groups past the end of the list keep running with an empty range: the shuffles below are warp-wide)
    extern __shared__ __align__(16) uint32_t qz_smem[];
    uint32_t* sr = qz_smem + (size_t)((threadIdx.x >> 5) * 4 + grp) * RS;
    unsigned char* sb = reinterpret_cast<unsigned char*>(sr);
    // element e of chunk i (dimension 32 i + 4 l8 + e) -> byte rec_pos: 64 (i >> 1) + 16 e + 8 (i & 1) + 4 (l8 & 1) + (l8 >> 1)
    const int lane_pos = 4 * (l8 & 1) + (l8 >> 1);
    float4 cv[W32T];
    uint32_t c_held = 0xffffffffu;
    uint2 my_item = make_uint2(0u, 0u);
    float my_ycd = 0.f;
    uint32_t my_wb = 0u, my_skip = 1u;
    if (i_begin + l8 < i_end) {  // lane l8 fetches what slot i_begin + l8 needs besides the two rows (pch <= 8)
        my_item = __ldg(&cl_items[i_begin + l8]);
        my_ycd = probe_dist[my_item.x];
        my_wb = q_wbase[my_item.x / (uint32_t)P] + slot_local[my_item.x];
        my_skip = (offsets && offsets[my_item.y + 1] == offsets[my_item.y]) ? 1u : 0u;
    }
    for (int sl = 0; sl < pch; sl++) {
        const size_t slot = i_begin + sl;
        const bool live = slot < i_end;                         // uniform inside the group, not inside the warp
        const uint32_t gw = __shfl_sync(FULL, my_item.x, sl, 8);
        const uint32_t c = __shfl_sync(FULL, my_item.y, sl, 8);
        const float ycd = __shfl_sync(FULL, my_ycd, sl, 8);
        const uint32_t wb = __shfl_sync(FULL, my_wb, sl, 8);
        const bool skip = __shfl_sync(FULL, my_skip, sl, 8) != 0u || !live;
        const uint32_t q = gw / (uint32_t)P;
        float4 rr[W32T];
        float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
        if (!skip) {
            const float4* yr = reinterpret_cast<const float4*>(y + (size_t)q * D);
            const float4* cr = reinterpret_cast<const float4*>(cent + (size_t)c * D);
            if (c != c_held) {
#pragma unroll
                for (int i = 0; i < W32T; i++) cv[i] = __ldg(&cr[8 * i + l8]);
                c_held = c;
            }
#pragma unroll
            for (int i = 0; i < W32T; i++) {
                const float4 yv = __ldg(&yr[8 * i + l8]);
                rr[i] = make_float4(__fsub_rn(yv.x, cv[i].x), __fsub_rn(yv.y, cv[i].y), __fsub_rn(yv.z, cv[i].z), __fsub_rn(yv.w, cv[i].w));
                mn = fminf(fminf(fminf(mn, rr[i].x), fminf(rr[i].y, rr[i].z)), rr[i].w);
                mx = fmaxf(fmaxf(fmaxf(mx, rr[i].x), fmaxf(rr[i].y, rr[i].z)), rr[i].w);
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
        }
        const float delta = __fmul_rn(__fsub_rn(mx, mn), SCALAR_1_15);  // rabitq.rs:307
        const float inv = __fdiv_rn(1.0f, delta);                       // :308 recip()
        int sum = 0;
        if (!skip) {
#pragma unroll
            for (int i = 0; i < W32T; i++) {
                const float rv[4] = {rr[i].x, rr[i].y, rr[i].z, rr[i].w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    int qi;
                    if (bias) {  // scalar_quantize_raw (src/utils.rs:194-209): truncate, saturate, NaN -> 0
                        const float f = __fadd_rn(__fmul_rn(__fsub_rn(rv[e], mn), inv), __ldg(&bias[32 * i + 4 * l8 + e]));
                        qi = (f != f) ? 0 : (f <= 0.0f ? 0 : (f >= 255.0f ? 255 : (int)f));
                    } else {
                        qi = cvtps_epi32(__fmul_rn(__fsub_rn(rv[e], mn), inv));
                    }
                    sum += qi;  // i32 lanes wrap like _mm256_add_epi32
                    sb[64 * (i >> 1) + 16 * e + 8 * (i & 1) + lane_pos] = (unsigned char)((qi & 15) << (3 - e));
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        if (!skip && l8 == 0) {
            float* rf = reinterpret_cast<float*>(sr + D / 4);
            rf[0] = mn;
            rf[1] = delta;
            rf[2] = __uint2float_rn((uint32_t)sum);  // `scalar_sum as f32`, rabitq.rs:322
            rf[3] = ycd;
            rf[4] = __fsqrt_rn(ycd);                 // rabitq.rs:346
            sr[D / 4 + 5] = (uint32_t)sum;
            sr[D / 4 + 6] = wb;
            sr[D / 4 + 7] = c;
        }
        __syncwarp();
        if (!skip) {
            uint4* rec = reinterpret_cast<uint4*>(qrec + slot * (size_t)rec_stride);
            for (int i = l8; i < RS / 4; i += 8) rec[i] = reinterpret_cast<const uint4*>(sr)[i];
        }
        __syncwarp();  // the staging area is rewritten by the next record
    }
}

// ---------------------------------------------------------------------------------------------------------
// Inverted probe lists for one round of probe ranks [p_lo, p_hi): cluster -> the (q, p) items probing it, and
// the scan work list (cluster, chunk of VT vectors).
// Rounds are expressed in EFFECTIVE ranks pe = p - p0[q] (p0 = first probe rank with vectors on this shard), so that the
// first round always sees the first candidates a query really visits here, also on a shard that does not own its
// nearest cluster.
// (`offsets` non-NULL on a shard: clusters held by another shard get no items, so K3 and K4 never see them.)
__global__ void bucket_count_kernel(const uint32_t* __restrict__ probe_ids, const uint32_t* __restrict__ q_p0, const uint32_t* __restrict__ offsets,
                                    size_t nq, int P, int p_lo, int p_hi, uint32_t* __restrict__ cl_count) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    int R = p_hi - p_lo;
    if (i >= nq * (size_t)R) return;
    size_t q = i / R;
    int p = p_lo + (int)(i % R) + (int)q_p0[q];
    if (p >= P) return;
    const uint32_t c = probe_ids[q * P + p];
    if (offsets && offsets[c + 1] == offsets[c]) return;
    atomicAdd(&cl_count[c], 1u);
}

// Work items of one cluster in one round: (chunks of VT vectors in the window) x (slices of at most MS of the m records
// probing the cluster).  Slicing the records keeps the items of a hot cluster (many queries, e.g. on a shard that serves
// the whole batch with few clusters) as small as everybody else's, so the persistent grid stays balanced.
RQ_DEV uint32_t chunk_count(uint32_t m, uint32_t n_c, int VT, uint32_t ch_min, uint32_t ch_max) {
    if (!m) return 0;
    const uint32_t nch = min((n_c + VT - 1) / VT, ch_max);
    return nch > ch_min ? nch - ch_min : 0;
}
RQ_DEV uint32_t slice_count(uint32_t m, uint32_t MS) { return (m + MS - 1) / MS; }

// One unit of scan work = (cluster, chunk of 128 vectors, slice of <= MS of the records probing the cluster), flattened so that
// the scan's producer warp needs ONE load to know everything about it.
struct __align__(16) ScanItem {
    uint32_t rec_begin;  // first record: index into cl_items
    uint32_t nr;         // records of this item
    uint32_t chunk_g;    // chunk index in the scan-layout copy of the codes (scan_layout_kernel)
    uint32_t chunk;      // chunk inside the cluster (visit position)
    uint32_t jbase;      // sorted position of the chunk's first vector
    uint32_t nv;         // vectors in the chunk (<= 128)
    uint32_t pad0, pad1;
};

__global__ void __launch_bounds__(1024) bucket_scan_kernel(const uint32_t* __restrict__ cl_count, const uint32_t* __restrict__ offsets,
                                                           int K, int VT, uint32_t MS, uint32_t ch_min, uint32_t ch_max, uint32_t* __restrict__ cl_start,
                                                           uint32_t* __restrict__ item_start, uint32_t* __restrict__ cl_cursor,
                                                           uint32_t* __restrict__ work_ctl /* [0] counter, [1] n_work */,
                                                           const uint32_t* __restrict__ spec_fail = nullptr) {
    __shared__ uint32_t wa[33], wb[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (K + 1023) / 1024;
    const int lo = min(K, tid * per), hi = min(K, lo + per);
    uint32_t sa = 0, sb = 0;
    for (int c = lo; c < hi; c++) {
        uint32_t m = cl_count[c], n_c = offsets[c + 1] - offsets[c];
        sa += m;
        sb += chunk_count(m, n_c, VT, ch_min, ch_max) * slice_count(m, MS);
    }
    uint32_t ia = sa, ib = sb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(FULL, ia, o), z = __shfl_up_sync(FULL, ib, o);
        if (lane >= o) { ia += x; ib += z; }
    }
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    if (warp == 0) {
        uint32_t a = wa[lane], b = wb[lane], xa = a, xb = b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t x = __shfl_up_sync(FULL, xa, o), z = __shfl_up_sync(FULL, xb, o);
            if (lane >= o) { xa += x; xb += z; }
        }
        wa[lane] = xa - a; wb[lane] = xb - b;
        if (lane == 31) { wa[32] = xa; wb[32] = xb; }
    }
    __syncthreads();
    uint32_t ra = wa[warp] + ia - sa, rb = wb[warp] + ib - sb;
    for (int c = lo; c < hi; c++) {
        uint32_t m = cl_count[c], n_c = offsets[c + 1] - offsets[c];
        cl_start[c] = ra;
        item_start[c] = rb;
        cl_cursor[c] = 0;
        ra += m;
        rb += chunk_count(m, n_c, VT, ch_min, ch_max) * slice_count(m, MS);
    }
    if (tid == 0) {
        cl_start[K] = wa[32];
        item_start[K] = wb[32];
        work_ctl[0] = 0;
        work_ctl[1] = (spec_fail && *spec_fail) ? 0u : wb[32];
    }
}

__global__ void bucket_fill_kernel(const uint32_t* __restrict__ probe_ids, const uint32_t* __restrict__ q_p0, const uint32_t* __restrict__ offsets,
                                   size_t nq, int P, int p_lo, int p_hi, const uint32_t* __restrict__ cl_start, uint32_t* __restrict__ cl_cursor,
                                   uint2* __restrict__ cl_items /* (q*P+p, cluster) */) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    int R = p_hi - p_lo;
    if (i >= nq * (size_t)R) return;
    size_t q = i / R;
    int p = p_lo + (int)(i % R) + (int)q_p0[q];
    if (p >= P) return;
    uint32_t c = probe_ids[q * P + p];
    if (offsets && offsets[c + 1] == offsets[c]) return;
    uint32_t pos = atomicAdd(&cl_cursor[c], 1u);
    cl_items[cl_start[c] + pos] = make_uint2((uint32_t)(q * P + p), c);
}

__global__ void work_items_kernel(const uint32_t* __restrict__ item_start, const uint32_t* __restrict__ cl_count,
                                  const uint32_t* __restrict__ cl_start, const uint32_t* __restrict__ offsets,
                                  const uint32_t* __restrict__ chunk_start, int K, uint32_t MS, uint32_t ch_min, ScanItem* __restrict__ work,
                                  const uint32_t* __restrict__ spec_fail = nullptr) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K || (spec_fail && *spec_fail)) return;
    const uint32_t s = item_start[c], e = item_start[c + 1];
    if (s == e) return;
    const uint32_t m = cl_count[c], nsl = slice_count(m, MS), off = offsets[c], n_c = offsets[c + 1] - off;
    for (uint32_t i = s; i < e; i++) {
        const uint32_t local = i - s, chunk = ch_min + local / nsl, slice = local % nsl;
        ScanItem w;
        w.rec_begin = cl_start[c] + slice * MS;
        w.nr = min(MS, m - slice * MS);
        w.chunk_g = chunk_start[c] + chunk;
        w.chunk = chunk;
        w.jbase = off + chunk * 128u;
        w.nv = min(128u, n_c - chunk * 128u);
        w.pad0 = w.pad1 = 0u;
        work[i] = w;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K4: the code scan.  calculate_rough_distance + asymmetric_binary_dot_product + binary_dot_product
// (src/rabitq.rs:336-367, src/utils.rs:113-135, src/simd.rs:326-384) fused with the threshold filter of
// HeapReRanker::rank_batch (src/rerank.rs:84) and an ORDER-PRESERVING compaction.
//
// The reference evaluates  abdp = sum_p popc(x & plane_p) << p,  which is the integer  sum_d bit_d * (q_u[d] & 15):  a
// (vectors x D) . (D x records) contraction of a 0/1 matrix with a 4-bit matrix.  Cluster-major, on the tensor cores, as a
// warp-specialised persistent kernel:
//   * the PRODUCER warp pulls work items (cluster, chunk of 128 vectors, <= 8*NT*sub records probing the cluster) off a global
//     counter and fills a ring of shared-memory stages with TMA bulk copies (cp.async.bulk, completion on an mbarrier): the
//     chunk's packed codes and Factors, kept in HBM in exactly the image the consumers read (scan_layout_kernel, built once per
//     index), and one copy per record (K3 wrote them in fragment order); it also resolves each record's threshold and round
//     window.  All global-memory latency -- work list, inverted list, records, codes -- lives in this warp, S-1 stages ahead;
//   * four CONSUMER warps own 32 vectors each (two m16 tiles) x 8*NT records (NT n8 tiles) of s32 accumulators.  Per k-step
//     (32 dimensions = one code word) a lane builds its A fragments from the packed word with one AND per register (+ one
//     shift for the upper nibbles) -- the bits are never expanded in memory -- loads its B fragments (query code bytes) with
//     one 128-bit shared-memory load per two k-steps and issues mma.sync.m16n8k32.s32.u8.u8.s32 (SASS IMMA.16832.U8.U8):
//     integer, exact;
//   * epilogue on the accumulator fragments: the estimator in the reference's association with IEEE roundings, the strict test
//     `rough < thr[q]`, one ballot per fragment register; a record's 32-vector bitmap word is assembled from four ballots by
//     rotate+mask, survivors' (rough, j) are packed at the start of the word's 32-entry block in vector order.  No atomics,
//     deterministic layout, visit order preserved.
// Row r of M-tile mt is vector 4*(r & 7) + 2*mt + (r >> 3) of the warp's 32: a lane's four accumulator rows are four
// CONSECUTIVE vectors (one 64-bit load per vector and k-step pair), and ballot bit 4g+t of register-row s lands on bitmap
// bit 4g+s by a rotation.
RQ_DEV void mma_u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    // volatile: a pure asm may be speculated out of the `nt < ntiles` guards (it was: every record tile was multiplied, used or not)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

RQ_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
RQ_DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
RQ_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {  // no arrive: only raises the pending transaction count
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
RQ_DEV void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
RQ_DEV void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
RQ_DEV void prefetch_l2_bulk(const void* src_gmem, uint32_t bytes) {  // no destination: warms L2 ahead of the TMA row gather
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
template <int SLEEP_NS = 0>
RQ_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends the warp in hardware only briefly; a waiter that expects to wait long (the scan's producer, which runs
    // stages ahead) sleeps between attempts instead of burning issue slots.  A lost transaction must fail loudly, never hang
    // the GPU: trap after ~4 s.
    const uint32_t addr = smem_u32(bar);
    auto attempt = [&]() -> uint32_t {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
        return ok;
    };
    if (attempt()) return;
    if constexpr (SLEEP_NS > 0) {
#pragma unroll 1
        for (uint32_t spin = 0; spin < 4000000000u / SLEEP_NS; spin++) {
            __nanosleep(SLEEP_NS);
            if (attempt()) return;
        }
    } else {
        const long long t0 = clock64();
#pragma unroll 1
        for (uint32_t spin = 1;; spin++) {
            if (attempt()) return;
            if ((spin & 1023u) == 0u && clock64() - t0 > 8000000000ll) break;
        }
    }
    __trap();
}

// Scan-layout copy of the codes and Factors, built once per index: for every cluster, chunks of 128 vectors (the last one
// zero-padded); a chunk's image is [D/64][128] uint2 -- words (2jj, 2jj+1) of the vector in slot (v & ~31) | (v & 3) << 3 |
// (v >> 2) & 7 -- followed, in a second array, by its 128 Factors in the same slot order.  One TMA bulk copy stages it.
__global__ void __launch_bounds__(128) scan_layout_kernel(const uint32_t* __restrict__ codes, const float4* __restrict__ factors,
                                                          const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ chunk_start,
                                                          int K, int JJ, uint2* __restrict__ scan_codes, float4* __restrict__ scan_fac) {
    const uint32_t b = blockIdx.x;
    int lo = 0, hi = K;  // largest c with chunk_start[c] <= b  (clusters without vectors own no chunk)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_start[mid] <= b) lo = mid; else hi = mid;
    }
    const uint32_t c = (uint32_t)lo, off = offsets[c], n_c = offsets[c + 1] - off;
    const uint32_t v0 = (b - chunk_start[c]) * 128u, nv = min(128u, n_c - v0);
    const uint2* src = reinterpret_cast<const uint2*>(codes + (size_t)(off + v0) * (2 * JJ));
    uint2* dst = scan_codes + (size_t)b * JJ * 128;
    const uint32_t tot = nv * (uint32_t)JJ;
    for (uint32_t i = threadIdx.x; i < 128u * JJ; i += 128u) {
        const uint32_t v = i / JJ, jj = i - v * JJ;
        dst[jj * 128 + ((v & ~31u) | ((v & 3u) << 3) | ((v >> 2) & 7u))] = i < tot ? src[i] : make_uint2(0u, 0u);
    }
    const uint32_t t = threadIdx.x;
    // padding slots of the last chunk: center_distance_square = +inf, so their estimate is +inf / NaN and `rough < thr` is false
    scan_fac[(size_t)b * 128 + ((t & ~31u) | ((t & 3u) << 3) | ((t >> 2) & 7u))] =
        t < nv ? factors[off + v0 + t] : make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
}

struct ScanArgs {
    const uint2* scan_codes;    // chunks x [D/64][128]
    const float4* scan_fac;     // chunks x 128
    const uint2* cl_items;      // (q*P+p, cluster) per cluster
    const ScanItem* work;
    uint32_t* work_ctl;         // [0] counter [1] n_work
    const unsigned char* qrec;  // this round's records (K3), list order, rec_pitch bytes apart
    const float* thr;           // per query
    const uint32_t* q_p0;       // per query: first probe rank with vectors on this shard
    uint32_t* bitmap;           // per slot word
    float2* entries;            // per slot word: 32 x (rough, j as bits)
    unsigned long long* counters;  // [0] survivors
    int P;
    int D;
    int rec_pitch;              // bytes between records in shared memory (== 64 mod 128: conflict-free fragment loads)
    int stages;                 // ring depth
    int sub;                    // passes of 8*NT records per stage
    uint32_t MS;                // records per work item (<= 8*NT*sub)
    // this round = visit positions (probe rank, 128-vector chunk) in [lo, hi), lexicographic
    int p_lo, ch_lo, p_hi, ch_hi;
};

constexpr int SCAN_CONSUMERS = 128;               // four consumer warps = 128 vectors per chunk
constexpr int SCAN_THREADS = SCAN_CONSUMERS;      // (vectors per chunk; the work-list builders use this name)
constexpr int SCAN_BLOCK = SCAN_CONSUMERS + 32;   // + the producer warp

__host__ __device__ __forceinline__ int scan_rec_pitch(int D) { return (D & 127) ? D + 128 : D + 64; }
// one stage: codes | Factors | records | (thr, flag) per record | header
__host__ __device__ __forceinline__ size_t scan_stage_bytes(int D, int nrs) {
    return (((size_t)16 * D + 2048 + (size_t)nrs * scan_rec_pitch(D) + (size_t)nrs * 8 + 32) + 127) & ~(size_t)127;
}
__host__ __device__ __forceinline__ size_t scan_smem_bytes(int D, int nrs, int stages) { return 128 + (size_t)stages * scan_stage_bytes(D, nrs); }

template <int NT, bool DENSE, int MINB>  // MINB: CTAs per SM the register allocation must allow (3 for small dims, 2 when shared memory allows no more anyway)
__global__ void __launch_bounds__(SCAN_BLOCK, MINB) scan_mma_kernel(ScanArgs a) {
    constexpr int NR = 8 * NT;  // records per pass
    extern __shared__ __align__(128) unsigned char scan_smem[];
    const int D = a.D, JJ = D >> 6, pitch = a.rec_pitch, S = a.stages, NRS = NR * a.sub;
    const size_t stage_bytes = scan_stage_bytes(D, NRS);
    uint64_t* full = reinterpret_cast<uint64_t*>(scan_smem);  // [S] producer -> consumers (32 arrivals + bytes)
    uint64_t* empty = full + 8;                               // [S] consumers -> producer (4 arrivals)
    unsigned char* stage0 = scan_smem + 128;
    const size_t o_fac = (size_t)16 * D, o_rec = o_fac + 2048, o_tf = o_rec + (size_t)NRS * pitch, o_hdr = o_tf + (size_t)NRS * 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 32); mbar_init(&empty[s], SCAN_CONSUMERS / 32); }
    }
    __syncthreads();

    if (warp == SCAN_CONSUMERS / 32) {
        // ---------------- producer ----------------------------------------------------------------------------------------
        const uint32_t n_work = a.work_ctl[1];
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&a.work_ctl[0], 1u);
        item = __shfl_sync(FULL, item, 0);
        for (uint32_t it = 0;; it++) {
            const int st = (int)(it % (uint32_t)S);
            const uint32_t ph = (it / (uint32_t)S) & 1u;
            unsigned char* sp = stage0 + (size_t)st * stage_bytes;
            int* hdr = reinterpret_cast<int*>(sp + o_hdr);
            const bool done = item >= n_work;
            ScanItem w;
            uint32_t next = 0;
            if (!done) {
                const uint4* wp = reinterpret_cast<const uint4*>(a.work + item);
                const uint4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
                w.rec_begin = w0.x; w.nr = w0.y; w.chunk_g = w0.z; w.chunk = w0.w; w.jbase = w1.x; w.nv = w1.y;
                if (lane == 0) next = atomicAdd(&a.work_ctl[0], 1u);  // in flight while this item is staged
            }
            mbar_wait<1000>(&empty[st], ph ^ 1u);  // the consumers released this stage (passes at once on its first use)
            if (done) {
                if (lane == 0) hdr[0] = -1;
                mbar_arrive(&full[st]);
                break;
            }
            if (lane == 0) {
                mbar_expect_tx(&full[st], (uint32_t)(16 * D + 2048));
                tma_bulk_g2s(sp, a.scan_codes + (size_t)w.chunk_g * JJ * 128, (uint32_t)(16 * D), &full[st]);
                tma_bulk_g2s(sp + o_fac, a.scan_fac + (size_t)w.chunk_g * 128, 2048u, &full[st]);
                hdr[0] = (int)w.nr; hdr[1] = (int)w.jbase; hdr[2] = (int)w.nv; hdr[3] = (int)w.chunk;
            }
            if (lane == 1) {  // the item's records are contiguous in HBM (K3 wrote them in list order, at the shared-memory pitch)
                const uint32_t bytes = w.nr * (uint32_t)pitch;
                mbar_expect_tx(&full[st], bytes);
                tma_bulk_g2s(sp + o_rec, a.qrec + (size_t)w.rec_begin * pitch, bytes, &full[st]);
            }
            float2* tf = reinterpret_cast<float2*>(sp + o_tf);
            for (uint32_t r = lane; r < (uint32_t)NRS; r += 32) {
                float th = __int_as_float(0xff800000);  // -inf: nothing passes (unused column, or a record outside this round's window)
                uint32_t fl = 0u;
                if (r < w.nr) {
                    const uint32_t id = __ldg(&a.cl_items[w.rec_begin + r]).x, q = id / (uint32_t)a.P;
                    const float thq = a.thr[q];
                    // is (rank, chunk) inside this round's window?
                    const int pr = (int)(id - q * (uint32_t)a.P) - (int)a.q_p0[q], ch = (int)w.chunk;
                    const bool ge_lo = pr > a.p_lo || (pr == a.p_lo && ch >= a.ch_lo);
                    const bool lt_hi = pr < a.p_hi || (pr == a.p_hi && ch < a.ch_hi);
                    fl = (ge_lo && lt_hi) ? 1u : 0u;
                    if (fl) th = thq;
                }
                tf[r] = make_float2(th, __uint_as_float(fl));
            }
            mbar_arrive(&full[st]);  // release: this lane's stores are visible to whoever observes the phase
            item = __shfl_sync(FULL, next, 0);
        }
        return;
    }

    // ---------------- consumers -------------------------------------------------------------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const uint32_t amask = 0x01010101u << t;
    uint32_t n_mine = 0;  // survivors counted by this lane (bitmap words it wrote)
    uint32_t lt[4], rot[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        lt[s] = (1u << (4 * g + s)) - 1u;  // vectors before 4g+s in the warp's 32
        rot[s] = (uint32_t)(t - s) & 31u;
    }
    for (uint32_t it = 0;; it++) {
        const int st = (int)(it % (uint32_t)S);
        const uint32_t ph = (it / (uint32_t)S) & 1u;
        unsigned char* sp = stage0 + (size_t)st * stage_bytes;
        mbar_wait<200>(&full[st], ph);
        const int4 hdr = *reinterpret_cast<const int4*>(sp + o_hdr);
        if (hdr.x < 0) break;
        const int nr_all = hdr.x;
        const uint32_t nv = (uint32_t)hdr.z, wloc = (uint32_t)hdr.w * (SCAN_CONSUMERS / 32) + warp;  // this warp's 32-vector word inside the cluster
        if ((uint32_t)(warp * 32) < nv) {
            const uint2* cp = reinterpret_cast<const uint2*>(sp) + warp * 32 + g;
            const float4* s_fac = reinterpret_cast<const float4*>(sp + o_fac);
            const float2* s_tf = reinterpret_cast<const float2*>(sp + o_tf);
            const uint32_t jbase = (uint32_t)hdr.y + warp * 32 + 4 * g;
#pragma unroll 1
            for (int r0 = 0; r0 < nr_all; r0 += NR) {
                const int nr = min(NR, nr_all - r0), ntiles = (nr + 7) >> 3;
                const unsigned char* s_rec = sp + o_rec + (size_t)r0 * pitch;
                int acc[2][NT][4];
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
#pragma unroll
                        for (int r = 0; r < 4; r++) acc[mt][nt][r] = 0;
                {
                    const unsigned char* bp = s_rec + (size_t)g * pitch + 16 * t;
#pragma unroll 1
                    for (int jj = 0; jj < JJ; jj++) {
                        uint32_t af[2][2][4];  // [k-step of the pair][M-tile][fragment register]
#pragma unroll
                        for (int s = 0; s < 4; s++) {  // vector 4g+s = M-tile s>>1, row g + 8*(s&1): registers (s&1) (k 0..15) and (s&1)+2 (k 16..31)
                            const uint2 w = cp[jj * 128 + 8 * s];
                            af[0][s >> 1][s & 1] = w.x & amask;
                            af[0][s >> 1][(s & 1) + 2] = (w.x >> 4) & amask;
                            af[1][s >> 1][s & 1] = w.y & amask;
                            af[1][s >> 1][(s & 1) + 2] = (w.y >> 4) & amask;
                        }
                        uint4 bf[NT];  // all B fragments first, then the MMAs k-step by k-step: 2*NT independent accumulators between dependent issues
#pragma unroll
                        for (int nt = 0; nt < NT; nt++)
                            if (nt < ntiles) bf[nt] = *reinterpret_cast<const uint4*>(bp + (size_t)nt * 8 * pitch + 64 * jj);
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            if (nt < ntiles) {  // warp-uniform
                                mma_u8(acc[0][nt], af[0][0], bf[nt].x, bf[nt].y);
                                mma_u8(acc[1][nt], af[0][1], bf[nt].x, bf[nt].y);
                            }
                        }
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            if (nt < ntiles) {
                                mma_u8(acc[0][nt], af[1][0], bf[nt].z, bf[nt].w);
                                mma_u8(acc[1][nt], af[1][1], bf[nt].z, bf[nt].w);
                            }
                        }
                    }
                }

                // ---- epilogue: estimator + filter + order-preserving compaction on the accumulator fragments ------------
                float4 fac[4];
#pragma unroll
                for (int s = 0; s < 4; s++) fac[s] = s_fac[warp * 32 + 8 * s + g];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    if (nt >= ntiles) continue;  // warp-uniform
                    uint32_t bal[2][4];
                    bool pass[2][4];
                    float rough[2][4];
                    float2* ebase[2];
                    uint32_t* bword[2];
                    uint32_t fl[2];
#pragma unroll
                    for (int o = 0; o < 2; o++) {
                        const int col = nt * 8 + 2 * t + o;
                        const float4 m0 = *reinterpret_cast<const float4*>(s_rec + (size_t)col * pitch + D);       // lo, delta, sum, ycd
                        const float4 m1 = *reinterpret_cast<const float4*>(s_rec + (size_t)col * pitch + D + 16);  // sqrt(ycd), -, wbase, -
                        const float2 tf = s_tf[r0 + col];  // threshold (-inf: column unused / outside the round's window), flag
                        fl[o] = __float_as_uint(tf.y);
                        const size_t word = (size_t)__float_as_uint(m1.z) + wloc;
                        ebase[o] = a.entries + word * 32;
                        bword[o] = a.bitmap + word;
#pragma unroll
                        for (int s = 0; s < 4; s++) {
                            const int iacc = acc[s >> 1][nt][(s & 1) * 2 + o];  // = 8 * abdp
                            // rabitq.rs:352-363, left-to-right:  ((((cds + ycd) + lo*ppc) + ((2*abdp - sum) * ip) * delta) - err * sqrt(ycd))
                            // 2*abdp is exact, so fma(0.25, 8*abdp, -sum) rounds once exactly like `2.0 * abdp as f32 - sum`
                            const float t1 = __fadd_rn(fac[s].w, m0.w);
                            const float t3 = __fadd_rn(t1, __fmul_rn(m0.x, fac[s].y));
                            const float t4 = fmaf(0.25f, __int2float_rn(iacc), -m0.z);
                            const float t6 = __fmul_rn(__fmul_rn(t4, fac[s].x), m0.y);
                            rough[o][s] = __fsub_rn(__fadd_rn(t3, t6), __fmul_rn(fac[s].z, m1.x));
                            if constexpr (DENSE) {
                                if ((uint32_t)(warp * 32 + 4 * g + s) < nv && fl[o])
                                    ebase[o][4 * g + s] = make_float2(rough[o][s], __uint_as_float((uint32_t)iacc >> 3));
                            } else {
                                // rerank.rs:84 (strict).  Padding vectors carry cds = +inf and switched-off columns thr = -inf: both fail here
                                pass[o][s] = rough[o][s] < tf.x;
                                bal[o][s] = __ballot_sync(FULL, pass[o][s]);
                            }
                        }
                    }
                    if constexpr (!DENSE) {
                        const uint32_t any = bal[0][0] | bal[0][1] | bal[0][2] | bal[0][3] | bal[1][0] | bal[1][1] | bal[1][2] | bal[1][3];
                        uint32_t bm[2] = {0u, 0u};
                        if (any) {  // warp-uniform
#pragma unroll
                            for (int o = 0; o < 2; o++)
#pragma unroll
                                for (int s = 0; s < 4; s++)  // ballot bit 4g'+t of register-row s  ->  bitmap bit 4g'+s
                                    bm[o] |= __funnelshift_r(bal[o][s], bal[o][s], rot[s]) & (0x11111111u << s);
#pragma unroll
                            for (int o = 0; o < 2; o++)
#pragma unroll
                                for (int s = 0; s < 4; s++)
                                    if (pass[o][s]) ebase[o][__popc(bm[o] & lt[s])] = make_float2(rough[o][s], __uint_as_float(jbase + s));
                        }
                        if (g == 0) {  // lanes 0..3 hold the flags / slot words of columns 2t, 2t+1
                            if (fl[0]) *bword[0] = bm[0];
                            if (fl[1]) *bword[1] = bm[1];
                            n_mine += __popc(bm[0]) + __popc(bm[1]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
    if constexpr (!DENSE) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_mine += __shfl_xor_sync(FULL, n_mine, o);
        if (lane == 0 && n_mine) atomicAdd(&a.counters[0], (unsigned long long)n_mine);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K5: exact rerank with the reference's SEQUENTIAL threshold semantics (HeapReRanker, src/rerank.rs:61-114).
//
// One warp per query streams the query's survivor words in visit order (probe rank ascending, position inside
// the cluster ascending) and queues the candidates whose estimate is still below the current threshold.  A full
// queue is a WAVE: the raw base rows of the wave are gathered into shared memory with one TMA bulk copy per row
// (cp.async.bulk, completion on an mbarrier), every exact squared L2 of the wave is computed in parallel --
// four threads per candidate, each owning two of the 8 AVX lanes of simd::l2_squared_distance (src/simd.rs:14-73) as one
// packed f32x2 chain, same per-lane order, same final reduction (l2_quad) -- and the wave is then replayed in order with the
// reference's two strict tests, one stretch (up to the next accepted candidate) per step.
// Heap contents, the threshold trajectory and the `precise` counter are exactly those of the sequential loop;
// speculation only costs extra gathers.
struct RerankArgs {
    const float* qpad;            // nq x D, unrotated padded queries (rabitq.rs:299)
    const float* base;            // n x D
    const uint32_t* map_ids;      // n
    const uint32_t* q_wbase;      // nq+1
    const uint32_t* slot_local;   // nq x P
    const uint32_t* q_p0;         // nq
    const uint32_t* bitmap;
    const float2* entries;
    float* heap_dist;             // nq x topk   state between rounds
    uint32_t* heap_ids;           // nq x topk
    uint32_t* heap_cnt;           // nq
    float* thr;                   // nq
    uint32_t* q_precise;          // nq (accumulated over rounds)
    float* h_recent;              // nq  HeuristicReRanker::recent_max_accurate between rounds
    uint32_t* h_wcount;           // nq  HeuristicReRanker::count between rounds
    unsigned long long* counters; // [1] exact computed (speculative) [2] precise
    float* out_dist;              // nq x topk (finalize)
    uint32_t* out_ids;
    uint32_t* out_count;
    int nq, P, D, topk;
    int R;                        // rows per wave (<= 32; rerank_cta_kernel: <= 8)
    int ns;                       // rerank_cta_kernel: row buffers in the ring
    const uint2* win;             // rerank_cta_kernel: this round's (first word, end word) per query (round_windows_kernel), or NULL
    // rerank_cta_kernel<.., SINK = 2> (distributed frozen round, source side): per query the survivor count of the round (upper bound
    // of the records; 0xffffffff = region overflow, flagged) and the offset of its records inside this shard's region of the home inbox
    const uint32_t* r2_cnt;
    const uint32_t* r2_off;
    size_t off_r2rec, off_r2tab;
    uint32_t cap2;
    int smem_per_warp;            // bytes
    int prefetch;                 // L2 prefetch of the survivors' base rows at discovery time (0 = off)
    uint32_t* dbg;                // NULL, or nq x 2 rounds x {cycles, waves, computed, enqueue (incl. the waves processed inside), wait, l2, replay, stage} (rabitq_debug_rerank_stats)
    // distributed record sink (template SINK != 0): records go straight into the inbox of the
    // query's HOME rank -- peer memory over NVLink (one process per GPU) or plain device memory (ranks in one process)
    unsigned char* const* peers;  // inbox base of every rank
    size_t off_r1cnt, off_r1rec;  // byte offsets inside an inbox (same layout on every rank)
    int world, rank, nq_local, r1cap;
};

// One reranked candidate as shipped to the query's home rank: everything HeapReRanker::rank_batch (src/rerank.rs:83-101)
// needs to replay its two strict tests (rough < thr, exact < thr) and the heap push, plus the probe rank that orders
// records of different shards in the reference's visit order.
struct __align__(16) SurvRec {
    float rough, exact;
    uint32_t id;   // original id (map_ids applied, rerank.rs:94)
    uint32_t p;    // probe rank of the cluster the candidate lives in
};

RQ_DEV void heap_recompute_max(const float* hd, int k, int lane, int& maxpos, float& thr) {
    if (k <= 32) {  // one redux + one ballot: the largest key, lowest slot among equals (what a sequential scan finds first)
        const uint32_t key = lane < k ? okey(hd[lane]) : 0u;
        const uint32_t bk = __reduce_max_sync(FULL, key);
        maxpos = __ffs(__ballot_sync(FULL, lane < k && key == bk)) - 1;
        thr = okey_to_float(bk);
        return;
    }
    uint32_t bk = 0;
    int bp = 0x7fffffff;
    for (int s = lane; s < k; s += 32) {
        uint32_t key = okey(hd[s]);
        if (key > bk || (key == bk && s < bp)) { bk = key; bp = s; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint32_t ok = __shfl_xor_sync(FULL, bk, o);
        int op = __shfl_xor_sync(FULL, bp, o);
        if (ok > bk || (ok == bk && op < bp)) { bk = ok; bp = op; }
    }
    maxpos = bp;
    thr = okey_to_float(bk);
}

// Exact squared L2 of one staged row against the query in the order of simd::l2_squared_distance (src/simd.rs:14-73): AVX lane l
// runs over elements l, l+8, ... ascending with diff = row - q rounded, then fma(diff, diff, acc); reduce_f32_256 adds the lanes
// as ((0+4)+(1+5))+((2+6)+(3+7)).  FOUR threads per candidate, thread v4 owns AVX lanes 2*v4 and 2*v4+1 as one packed f32x2
// (64-bit shared-memory loads, sub.rn.f32x2 + fma.rn.f32x2: half the instructions of a lane-per-thread loop).  The chain of D/8
// dependent FMAs is the critical path of a wave, so the operands of the next 4 steps are loaded while the current 4 are
// consumed (D is a multiple of 64).  All four threads return the same bits.
RQ_DEV float l2_quad(const float* __restrict__ row, const float* __restrict__ qv, int D, int v4) {
    const f32x2* rp = reinterpret_cast<const f32x2*>(row) + v4;  // pair (2 v4, 2 v4 + 1) of 8-element step i at rp[4 i]
    const f32x2* qp = reinterpret_cast<const f32x2*>(qv) + v4;
    f32x2 x[4], q[4], acc = 0ull;
#pragma unroll
    for (int i = 0; i < 4; i++) { x[i] = rp[4 * i]; q[i] = qp[4 * i]; }
    const int nblk = D / 32;
    for (int b = 0; b < nblk; b++) {
        const int o = 16 * min(b + 1, nblk - 1);  // the last block re-reads itself (unused)
        f32x2 nx[4], nq[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { nx[i] = rp[o + 4 * i]; nq[i] = qp[o + 4 * i]; }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const f32x2 f = sub2(x[i], q[i]);
            acc = fma2(f, f, acc);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) { x[i] = nx[i]; q[i] = nq[i]; }
    }
    float lo, hi, olo, ohi;
    unpack2(acc, lo, hi);
    unpack2(__shfl_xor_sync(FULL, acc, 2), olo, ohi);
    const float s = __fadd_rn(__fadd_rn(lo, olo), __fadd_rn(hi, ohi));  // v4 even: c0 + c1, odd: c2 + c3
    return __fadd_rn(s, __shfl_xor_sync(FULL, s, 1));
}

// Same distance with the ROW IN GLOBAL MEMORY (a rotated centroid, L2-resident) and the query in shared memory: 64-bit read-only
// loads, eight steps in flight.  Used by the prefilter's exact recheck (prefilter.cuh); diff = row - q as in the reference.
RQ_DEV float l2_quad_global(const float* __restrict__ row, const float* __restrict__ qs, int D, int v4) {
    const f32x2* rp = reinterpret_cast<const f32x2*>(row) + v4;
    const f32x2* qp = reinterpret_cast<const f32x2*>(qs) + v4;
    f32x2 acc = 0ull;
#pragma unroll 8
    for (int i = 0; i < D / 8; i++) {
        const f32x2 f = sub2(__ldg(&rp[4 * i]), qp[4 * i]);
        acc = fma2(f, f, acc);
    }
    float lo, hi, olo, ohi;
    unpack2(acc, lo, hi);
    unpack2(__shfl_xor_sync(FULL, acc, 2), olo, ohi);
    const float s = __fadd_rn(__fadd_rn(lo, olo), __fadd_rn(hi, ohi));
    return __fadd_rn(s, __shfl_xor_sync(FULL, s, 1));
}

// Latency-first variant for the rerank replay: EIGHT threads per candidate, thread l = AVX lane l of simd::l2_squared_distance as one
// scalar chain (dependent FFMA latency 4 cycles; the packed f32x2 chain of l2_quad issues at half rate and the wave's critical path
// was 2.4 k cycles at D = 960), NC candidates interleaved per thread for instruction-level parallelism.  Rows and query in shared
// memory; the 8 lanes of a group read 32 contiguous bytes per step.  Every lane of a group returns the same bits.
template <int NC>
RQ_DEV void l2_oct(const float* const (&row)[NC], const float* __restrict__ qv, int D, int l, float (&out)[NC]) {
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = 0.0f;
#pragma unroll 1
    for (int s0 = 0; s0 < D; s0 += 64) {  // 8 steps of 8 elements per iteration (D is a multiple of 64); (a register-prefetched form was measured slower)
        float x[NC][8], q[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            q[u] = qv[s0 + 8 * u + l];
#pragma unroll
            for (int c = 0; c < NC; c++) x[c][u] = row[c][s0 + 8 * u + l];
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const float d = __fsub_rn(x[c][u], q[u]);
                acc[c] = fmaf(d, d, acc[c]);
            }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {  // reduce_f32_256: c_i = s_i + s_{i+4}; (c0 + c1) + (c2 + c3)   (fp32 addition commutes: all lanes agree)
        float a = acc[c];
        a = __fadd_rn(a, __shfl_xor_sync(FULL, a, 4));
        a = __fadd_rn(a, __shfl_xor_sync(FULL, a, 1));
        out[c] = __fadd_rn(a, __shfl_xor_sync(FULL, a, 2));
    }
}

// HEUR = false: HeapReRanker (src/rerank.rs:61-114).  HEUR = true: HeuristicReRanker (src/rerank.rs:117-176): the filter
// threshold is the largest accepted distance of the last WINDOW_SIZE = 12 accepted candidates (src/consts.rs:12); every
// accepted candidate is a result candidate and get_result keeps the topk smallest, which is what the k-slot buffer holds.
// SINK = 0: single-GPU replay.  SINK = 1 (distributed round 1, run by the shard that owns the query's nearest non-empty
// cluster): the same sequential replay, and every candidate the reference computes an exact distance for is also written
// to the home rank's inbox.  (The distributed round 2 has no sequential dependency and is a flat kernel: r2_exact_kernel.)
template <bool HEUR, int SINK>
__global__ void __launch_bounds__(128, 5) rerank_kernel(RerankArgs a, int p_lo, int ch_lo, int p_hi, int ch_hi, int first, int finalize) {
    extern __shared__ __align__(16) unsigned char rr_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int q = blockIdx.x * wpb + warp;
    if (q >= a.nq) return;
    const int D = a.D, k = a.topk, R = a.R, pitch = D + 8;  // +8 floats: the 4 candidates of a group hit disjoint banks
    const uint32_t rowbytes = (uint32_t)D * 4u;
    unsigned char* wbase = rr_smem_raw + (size_t)warp * a.smem_per_warp;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase);  // two barriers, one per wave buffer
    float* qv = reinterpret_cast<float*>(wbase + 16);
    float* rows = qv + D;                                   // [2][R][pitch]
    float* hd = rows + (size_t)2 * R * pitch;
    uint32_t* hid = reinterpret_cast<uint32_t*>(hd + k);
    float* qr = reinterpret_cast<float*>(hid + k);          // [2][32]
    uint32_t* qj = reinterpret_cast<uint32_t*>(qr + 64);    // [2][32]
    float2* sent = reinterpret_cast<float2*>(qj + 64);      // [4][32] staged survivors
    uint32_t* stot = reinterpret_cast<uint32_t*>(sent + 128);  // [4]
    const uint32_t lt_mask = (1u << lane) - 1u;

    const long long dbg_t0 = a.dbg ? clock64() : 0ll;
    uint32_t dbg_waves = 0, dbg_wait = 0, dbg_l2 = 0, dbg_replay = 0, dbg_stage = 0, dbg_enq = 0;
    const bool dbg_on = a.dbg != nullptr;
    if (lane == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); }
    // prologue: every independent global load is issued before the first dependent use (the kernel is a chain of
    // memory latencies; the fewer links the better)
    int cnt = first ? 0 : (int)a.heap_cnt[q];
    float thr = first ? 3.402823466e+38f : a.thr[q];   // the filter threshold of the reranker
    float hmax = 3.402823466e+38f;                      // largest distance among the k kept results (when cnt == k)
    float recent = first ? -3.402823466e+38f : a.h_recent[q];
    uint32_t wcount = first ? 0u : a.h_wcount[q];
    const uint32_t wb = a.q_wbase[q], wend = a.q_wbase[q + 1];
    const int p0 = (int)a.q_p0[q];
    if constexpr (SINK == 0) {  // issued before the dependent window lookups below: one link less in the latency chain
        const float4* src = reinterpret_cast<const float4*>(a.qpad + (size_t)q * D);
        float4* dst = reinterpret_cast<float4*>(qv);
#pragma unroll 8
        for (int d = lane; d < D / 4; d += 32) dst[d] = __ldg(&src[d]);
    }
    // word window of this round: from (p_lo, ch_lo) to (p_hi, ch_hi) in visit order; a chunk is 128 vectors = 4 words
    auto word_at = [&](int pe, int ch) -> uint32_t {  // pe = effective rank
        const int p = pe + p0;
        if (p >= a.P) return wend;
        const uint32_t s0 = wb + a.slot_local[(size_t)q * a.P + p];
        const uint32_t s1 = (p + 1 < a.P) ? wb + a.slot_local[(size_t)q * a.P + p + 1] : wend;
        return min(s0 + (uint32_t)ch * (SCAN_THREADS / 32), s1);
    };
    const uint32_t wlo = word_at(p_lo, ch_lo), whi = word_at(p_hi, ch_hi);
    if (SINK != 0 && wlo < whi) {  // distributed: most queries have no candidates in this shard's round-1 window; skip the row-sized load
        const float4* src = reinterpret_cast<const float4*>(a.qpad + (size_t)q * D);
        float4* dst = reinterpret_cast<float4*>(qv);
#pragma unroll 8
        for (int d = lane; d < D / 4; d += 32) dst[d] = __ldg(&src[d]);
    }
    // record sinks
    const int home = SINK ? q / a.nq_local : 0, ql = SINK ? q - home * a.nq_local : 0;
    SurvRec* rdst = nullptr;
    uint32_t nrec = 0;
    if constexpr (SINK == 1)
        rdst = reinterpret_cast<SurvRec*>(a.peers[home] + a.off_r1rec) + ((size_t)a.rank * a.nq_local + ql) * a.r1cap;
    int maxpos = 0;
    for (int s = lane; s < cnt; s += 32) {
        hd[s] = a.heap_dist[(size_t)q * k + s];
        hid[s] = a.heap_ids[(size_t)q * k + s];
    }
    __syncwarp();
    if (cnt == k) {
        heap_recompute_max(hd, k, lane, maxpos, hmax);
        if constexpr (!HEUR) thr = hmax;
    }
    uint32_t precise = 0, computed = 0, par = 0;
    int fill = 0, f = 0, pend_n = 0;  // wave being filled, its fill level, size of the closed-but-unprocessed wave (buffer f^1)

    // process one closed wave: wait for its rows, all exact distances in parallel, in-order replay
    auto process = [&](int w, int n) {
        long long dt0 = dbg_on ? clock64() : 0ll;
        if (lane == 0) mbar_arrive(&mbar[w]);
        mbar_wait(&mbar[w], (par >> w) & 1u);
        par ^= 1u << w;
        if (dbg_on) { const long long t = clock64(); dbg_wait += (uint32_t)(t - dt0); dt0 = t; }
        const bool mine = lane < n;
        const float rough = mine ? qr[w * 32 + lane] : 0.0f;
        const uint32_t j = mine ? qj[w * 32 + lane] : 0u;
        const uint32_t act = __ballot_sync(FULL, mine && rough < thr);
        computed += n;
        dbg_waves++;
        if (act) {
            const float* rw = rows + (size_t)w * R * pitch;
            float exact = 0.0f;
            for (int g = 0; g < n; g += 8) {  // 8 candidates per pass: the 8 lanes of group c compute candidates g + c and g + 4 + c
                const int grp = lane >> 3, l8 = lane & 7;
                if (g + 4 < n) {
                    const float* const rp[2] = {rw + (size_t)min(g + grp, R - 1) * pitch, rw + (size_t)min(g + 4 + grp, R - 1) * pitch};
                    float d2[2];
                    l2_oct<2>(rp, qv, D, l8, d2);
                    const float r0 = __shfl_sync(FULL, d2[0], 8 * ((lane - g) & 3)), r1 = __shfl_sync(FULL, d2[1], 8 * ((lane - g) & 3));
                    if (lane >= g && lane < g + 4) exact = r0;
                    if (lane >= g + 4 && lane < g + 8) exact = r1;
                } else {
                    const float* const rp[1] = {rw + (size_t)min(g + grp, R - 1) * pitch};
                    float d2[1];
                    l2_oct<1>(rp, qv, D, l8, d2);
                    const float r0 = __shfl_sync(FULL, d2[0], 8 * ((lane - g) & 3));
                    if (lane >= g && lane < g + 4) exact = r0;
                }
            }
            if (dbg_on) { const long long t = clock64(); dbg_l2 += (uint32_t)(t - dt0); dt0 = t; }
            // in-order replay (rerank.rs:83-101).  The threshold only moves when a candidate is ACCEPTED (rough < thr and
            // exact < thr), so the stretch up to the next accepted candidate is evaluated in one step: every lane of the
            // stretch with rough < thr is a candidate the reference computes an exact distance for.
            uint32_t rem = act;
            uint32_t mid = 0;
            if constexpr (SINK == 1) mid = ((act >> lane) & 1u) ? a.map_ids[j] : 0u;
            while (rem) {
                const bool pass = ((rem >> lane) & 1u) && rough < thr;
                const uint32_t pm = __ballot_sync(FULL, pass);
                const uint32_t am = __ballot_sync(FULL, pass && exact < thr);
                const int t = __ffs(am) - 1;                                    // the next accepted candidate (-1: none)
                const uint32_t upto = am ? ((2u << t) - 1u) : FULL;             // lanes 0..t
                const uint32_t cm = pm & upto;
                precise += __popc(cm);
                if constexpr (SINK == 1) {
                    const uint32_t pos = nrec + __popc(cm & lt_mask);
                    if (((cm >> lane) & 1u) && pos < (uint32_t)a.r1cap) {
                        SurvRec rec;
                        rec.rough = rough; rec.exact = exact; rec.id = mid; rec.p = (uint32_t)(p0 + p_lo);
                        rdst[pos] = rec;
                    }
                    nrec += __popc(cm);
                }
                if (!am) break;
                rem &= ~upto;
                const float ex = __shfl_sync(FULL, exact, t);
                if (!HEUR || cnt < k || ex < hmax) {
                    const int slot = cnt < k ? cnt : maxpos;
                    if (lane == t) { hd[slot] = ex; hid[slot] = SINK == 1 ? mid : j; }  // SINK == 0: the POSITION; map_ids at finalize
                    if (cnt < k) cnt++;
                    __syncwarp();
                    if (cnt == k) {
                        heap_recompute_max(hd, k, lane, maxpos, hmax);
                        if constexpr (!HEUR) thr = hmax;  // rerank.rs:98-100
                    }
                }
                if constexpr (HEUR) {  // rerank.rs:155-162
                    wcount++;
                    recent = fmaxf(recent, ex);
                    if (wcount >= 12u) {
                        thr = recent;
                        wcount = 0;
                        recent = -3.402823466e+38f;
                    }
                }
            }
        }
        if (dbg_on) dbg_replay += (uint32_t)(clock64() - dt0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // rows of this buffer are rewritten by later bulk copies
        __syncwarp();
    };

    // queue candidates in visit order; the row gather of each is issued right here, so it is in flight while the
    // stream continues and while the previous wave is being replayed
    auto enqueue = [&](uint32_t pm, float rough, uint32_t j) {
        const long long de0 = dbg_on ? clock64() : 0ll;
        while (pm) {
            const int space = R - fill;
            const int rank = __popc(pm & lt_mask);
            const bool take = ((pm >> lane) & 1u) && rank < space;
            const uint32_t took = __ballot_sync(FULL, take);
            if (lane == __ffs(took) - 1) mbar_expect_tx(&mbar[f], rowbytes * (uint32_t)__popc(took));  // one update for the whole batch of rows
            __syncwarp();
            if (take) {
                const int slot = fill + rank;
                qr[f * 32 + slot] = rough;
                qj[f * 32 + slot] = j;
                tma_bulk_g2s(rows + ((size_t)f * R + slot) * pitch, a.base + (size_t)j * D, rowbytes, &mbar[f]);
            }
            pm &= ~took;
            fill += __popc(took);
            __syncwarp();
            if (fill == R) {
                if (pend_n) process(f ^ 1, pend_n);
                pend_n = R;
                f ^= 1;
                fill = 0;
            }
        }
        if (dbg_on) dbg_enq += (uint32_t)(clock64() - de0);
    };

    // Super-block of SB x 32 words: SB independent bitmap loads per lane, then the first 32 survivors of every block
    // loaded back to back and staged in shared memory; the (rolled, single-instance) processing loop follows.
    constexpr int SB = 4;
    uint32_t mnext[SB];  // bitmap words of the NEXT super-block, loaded one super-block ahead (one DRAM round trip less per block)
#pragma unroll
    for (int u = 0; u < SB; u++) {
        const uint32_t idx = wlo + u * 32 + lane;
        mnext[u] = idx < whi ? a.bitmap[idx] : 0u;
    }
    for (uint32_t w0 = wlo; w0 < whi; w0 += 32 * SB) {
        const long long ds0 = dbg_on ? clock64() : 0ll;
        {
            uint32_t m[SB], incl[SB];
#pragma unroll
            for (int u = 0; u < SB; u++) {
                m[u] = mnext[u];
                const uint32_t idx = w0 + 32 * SB + u * 32 + lane;
                mnext[u] = idx < whi ? a.bitmap[idx] : 0u;
            }
#pragma unroll
            for (int u = 0; u < SB; u++) {
                uint32_t x = __popc(m[u]);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(FULL, x, o);
                    if (lane >= o) x += t;
                }
                incl[u] = x;
            }
            float2 ent[SB];
#pragma unroll
            for (int u = 0; u < SB; u++) {
                const uint32_t T = __shfl_sync(FULL, incl[u], 31);
                int pos = 0;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    uint32_t pv = __shfl_sync(FULL, incl[u], pos + s - 1);
                    if (pv <= (uint32_t)lane) pos += s;
                }
                pos = min(pos, 31);
                const uint32_t src_excl = __shfl_sync(FULL, incl[u] - __popc(m[u]), pos);
                ent[u] = make_float2(3.402823466e+38f, 0.0f);
                if ((uint32_t)lane < T) ent[u] = a.entries[(size_t)(w0 + u * 32 + pos) * 32 + (lane - src_excl)];
                if (lane == 0) stot[u] = T;
            }
            // The gather of a row is a DRAM round trip in front of every wave; the survivors of the whole super-block are
            // known here, long before their wave is filled, so their rows are pulled into L2 now (while the threshold is
            // still +inf only as many as two waves hold).
            if (a.prefetch) {
#pragma unroll
                for (int u = 0; u < SB; u++)
                    if (ent[u].x < thr && (thr < 3.402823466e+38f || lane < 2 * R))
                        prefetch_l2_bulk(a.base + (size_t)__float_as_uint(ent[u].y) * D, rowbytes);
            }
#pragma unroll
            for (int u = 0; u < SB; u++) sent[u * 32 + lane] = ent[u];
            __syncwarp();
        }
        if (dbg_on) dbg_stage += (uint32_t)(clock64() - ds0);
        for (int u = 0; u < SB; u++) {
            const uint32_t T = stot[u];
            if (T == 0) continue;  // uniform
            {
                const float2 en = sent[u * 32 + lane];
                const uint32_t pm = __ballot_sync(FULL, (uint32_t)lane < T && en.x < thr);
                if (pm) enqueue(pm, en.x, __float_as_uint(en.y));
            }
            if (T > 32) {  // dense block (loose threshold, e.g. the first probed cluster): further chunks on demand
                const uint32_t idx = w0 + u * 32 + lane;
                const uint32_t mm = idx < whi ? a.bitmap[idx] : 0u;
                uint32_t inc = __popc(mm);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(FULL, inc, o);
                    if (lane >= o) inc += t;
                }
                const uint32_t exc = inc - __popc(mm);
                auto load_chunk = [&](uint32_t e0) -> float2 {  // survivors e0 .. e0+31 of the block, one per lane
                    const uint32_t e = e0 + lane;
                    int pos = 0;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) {
                        uint32_t pv = __shfl_sync(FULL, inc, pos + s - 1);
                        if (pv <= e) pos += s;
                    }
                    pos = min(pos, 31);
                    const uint32_t src_excl = __shfl_sync(FULL, exc, pos);
                    float2 en = make_float2(3.402823466e+38f, 0.0f);
                    if (e < T) en = a.entries[(size_t)(w0 + u * 32 + pos) * 32 + (e - src_excl)];
                    return en;
                };
                float2 en = load_chunk(32);
                for (uint32_t e0 = 32; e0 < T; e0 += 32) {  // the next chunk's entries fly while this chunk's waves are replayed
                    float2 nxt = make_float2(3.402823466e+38f, 0.0f);
                    if (e0 + 32 < T) nxt = load_chunk(e0 + 32);
                    const uint32_t pm = __ballot_sync(FULL, e0 + lane < T && en.x < thr);
                    if (pm) enqueue(pm, en.x, __float_as_uint(en.y));
                    en = nxt;
                }
            }
        }
        __syncwarp();
    }
    if (pend_n) process(f ^ 1, pend_n);
    if (fill) process(f, fill);

    if constexpr (SINK == 1) {  // only the owner of the window has words; it tells the home rank how many records to replay
        if (lane == 0 && wlo < whi)
            reinterpret_cast<uint32_t*>(a.peers[home] + a.off_r1cnt)[(size_t)a.rank * a.nq_local + ql] = min(nrec, (uint32_t)a.r1cap);
    }
    if (a.dbg && lane == 0) {
        uint32_t* o = a.dbg + ((size_t)q * 2 + (first ? 0 : 1)) * 8;
        o[0] = (uint32_t)(clock64() - dbg_t0); o[1] = dbg_waves; o[2] = computed; o[3] = dbg_enq;
        o[4] = dbg_wait; o[5] = dbg_l2; o[6] = dbg_replay; o[7] = dbg_stage;
    }
    if (lane == 0) {
        a.q_precise[q] = (first ? 0u : a.q_precise[q]) + precise;
        atomicAdd(&a.counters[1], (unsigned long long)computed);
        atomicAdd(&a.counters[2], (unsigned long long)precise);
    }
    if (!finalize) {
        for (int s = lane; s < cnt; s += 32) {
            a.heap_dist[(size_t)q * k + s] = hd[s];
            a.heap_ids[(size_t)q * k + s] = hid[s];
        }
        if (lane == 0) { a.heap_cnt[q] = (uint32_t)cnt; a.thr[q] = thr; a.h_recent[q] = recent; a.h_wcount[q] = wcount; }
    } else {
        // ascending by (distance, id): rank by counting.  The heap holds positions (no map_ids load inside the sequential replay);
        // the original ids (rerank.rs:94) are looked up here, k independent loads.
        __syncwarp();
        if constexpr (SINK == 0) {
            for (int s = lane; s < cnt; s += 32) hid[s] = a.map_ids[hid[s]];
            __syncwarp();
        }
        for (int s = lane; s < k; s += 32) {
            if (s < cnt) {
                const uint32_t ks = okey(hd[s]), is = hid[s];
                int rank = 0;
                for (int t = 0; t < cnt; t++) {
                    const uint32_t kt = okey(hd[t]), itt = hid[t];
                    rank += (kt < ks) || (kt == ks && (itt < is || (itt == is && t < s)));
                }
                a.out_dist[(size_t)q * k + rank] = hd[s];
                a.out_ids[(size_t)q * k + rank] = is;
            } else {
                a.out_dist[(size_t)q * k + s] = __int_as_float(0x7f800000);
                a.out_ids[(size_t)q * k + s] = 0xffffffffu;
            }
        }
        if (lane == 0) a.out_count[q] = (uint32_t)cnt;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K5 (single-GPU form): the same exact sequential replay as rerank_kernel, one CTA per query, warp-specialised.
//
// rerank_kernel gives a query ONE warp, which streams the survivors, issues the row gathers, computes the exact distances and
// replays the reference's tests, one after the other: a launch lasts as long as its slowest query's chain of latencies (7 % of
// the warp slots active).  Here the three jobs run concurrently on different warps of the query's CTA and
// meet through mbarriers:
//   * warp 0, the PRODUCER, streams the query's survivor words in visit order (bitmap words two blocks ahead, entries one block
//     ahead), keeps the candidates whose estimate is below the threshold as it stands (a stale threshold is only LARGER: the
//     queue is a superset of what the reference computes) and packs them into WAVES of R rows; every row is gathered with one TMA
//     bulk copy into one of two row buffers, completion on the wave's `full` barrier;
//   * warps 2.., the COMPUTE warps, turn a full wave into exact squared distances: eight lanes per candidate, lane l = AVX lane l
//     of simd::l2_squared_distance (src/simd.rs:14-73) as one scalar FMA chain (l2_oct), NC candidates per group, all groups of
//     all compute warps at once -- the dependent chain of D/8 FMAs is paid once per wave, not once per four rows;
//   * warp 1, the REPLAY warp, owns the k-slot result buffer and the threshold: it walks the wave in order with the reference's
//     two strict tests (src/rerank.rs:84,92), stretch by stretch, and publishes the threshold for the other two.
// Row buffers are double-buffered between producer and compute warps, the (rough, position, exact) wave records live in a ring
// of four between producer, compute and replay.  Heap contents, threshold trajectory and `precise` are exactly those of the
// sequential loop; speculation costs extra gathers only.
// Word windows [first, end) of every rerank round of every query in the survivor-slot space: the round bounds are visit positions
// (effective probe rank, 128-vector chunk); turning them into words is a chain of three dependent loads (q_p0 -> slot_local ->
// ...) that would otherwise open every query's replay -- done here, once, on the side stream.
struct RoundBounds { int n; int p[17]; int ch[17]; };  // n positions = n - 1 rounds
__global__ void round_windows_kernel(const uint32_t* __restrict__ q_wbase, const uint32_t* __restrict__ slot_local, const uint32_t* __restrict__ q_p0,
                                     int nq, int P, RoundBounds b, uint2* __restrict__ win /* (n - 1) x nq */,
                                     const uint32_t* __restrict__ spec_fail = nullptr) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    if (spec_fail && *spec_fail) {  // the batch is going to be repeated: nothing to replay
        for (int i = 1; i < b.n; i++) win[(size_t)(i - 1) * nq + q] = make_uint2(0u, 0u);
        return;
    }
    const uint32_t wb = q_wbase[q], wend = q_wbase[q + 1];
    const int p0 = (int)q_p0[q];
    uint32_t prev = 0;
    for (int i = 0; i < b.n; i++) {
        const int p = b.p[i] + p0;
        uint32_t w = wend;
        if (p < P) {
            const uint32_t s0 = wb + slot_local[(size_t)q * P + p];
            const uint32_t s1 = (p + 1 < P) ? wb + slot_local[(size_t)q * P + p + 1] : wend;
            w = min(s0 + (uint32_t)b.ch[i] * (SCAN_THREADS / 32), s1);
        }
        if (i > 0) win[(size_t)(i - 1) * nq + q] = make_uint2(prev, w);
        prev = w;
    }
}

#ifndef RQ_RC_SB1
#define RQ_RC_SB1 8
#endif
// blocks of 32 survivor words per super-block of the producer's stream: 8 for long rows (few queries per SM anyway), 2 for short rows
// (NC = 2), where registers and shared memory per CTA decide how many queries an SM keeps in flight
__host__ __device__ constexpr int rerank_cta_sb(int nc) { return nc == 2 ? 2 : RQ_RC_SB1; }
__host__ __device__ __forceinline__ size_t rerank_cta_smem(int D, int topk, int R, int ns, int nc) {
    const int sb = rerank_cta_sb(nc);
    return 528 + (size_t)D * 4 + (size_t)ns * R * (D + 8) * 4 + 2 * (size_t)topk * 4 + 5 * 16 * 8 * 4 + (size_t)sb * 32 * 16 + sb * 4 + 16;
}
// SINK = 1 (distributed round 1, run by the shard that owns the query's nearest non-empty cluster): the same replay, and every candidate
// the reference computes an exact distance for is also written to the home rank's inbox (peer memory over NVLink), as in rerank_kernel.
template <bool HEUR, int NC, int SINK = 0>
__global__ void __launch_bounds__(320) rerank_cta_kernel(RerankArgs a, int p_lo, int ch_lo, int p_hi, int ch_hi, int first, int finalize) {
    extern __shared__ __align__(16) unsigned char rc_smem_raw[];
    constexpr int RC_SB = rerank_cta_sb(NC);
    constexpr uint32_t NM = 16, RW = 8;  // wave records in the ring, candidates per record (R <= 4 NC <= 8)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ncw = (int)(blockDim.x >> 5) - 2;
    const int q = blockIdx.x;
    const int D = a.D, k = a.topk, R = a.R, pitch = D + 8;
    const uint32_t NS = (uint32_t)a.ns;  // row buffers in the ring (<= 8)
    const uint32_t rowbytes = (uint32_t)D * 4u;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(rc_smem_raw);  // [8]  rows of a wave landed (producer arrive + TMA bytes)
    uint64_t* bar_rfree = bar_full + 8;                             // [8]  the wave's compute warp is done with the row buffer
    uint64_t* bar_exact = bar_full + 16;                            // [16] the wave's exact distances are written
    uint64_t* bar_mfree = bar_full + 32;                            // [16] the replay warp is done with the wave record
    volatile float* thr_s = reinterpret_cast<volatile float*>(rc_smem_raw + 384);
    volatile uint32_t* wn = reinterpret_cast<volatile uint32_t*>(rc_smem_raw + 388);  // [16] candidates of the wave
    volatile uint32_t* wlast = wn + NM;                                              // [16] 1: last wave of the query, 2: terminator
    float* qv = reinterpret_cast<float*>(rc_smem_raw + 528);
    float* rows = qv + D;                                    // [NS][R][pitch]
    float* hd = rows + (size_t)NS * R * pitch;
    uint32_t* hid = reinterpret_cast<uint32_t*>(hd + k);
    float* qr = reinterpret_cast<float*>(hid + k);           // [NM][RW] rough
    uint32_t* qj = reinterpret_cast<uint32_t*>(qr + NM * RW);  // [NM][RW] position
    float* ex = reinterpret_cast<float*>(qj + NM * RW);      // [NM][RW] exact
    uint32_t* smid = reinterpret_cast<uint32_t*>(ex + NM * RW);  // [NM][RW] original id (SINK: looked up by the compute warps)
    uint32_t* sqw = smid + NM * RW;                              // [NM][RW] SINK = 2: slot word of the candidate -> its probe rank
    float2* s_en = reinterpret_cast<float2*>(sqw + NM * RW);   // [RC_SB][32] the producer's stash of one super-block's first entries
    uint32_t* s_incl = reinterpret_cast<uint32_t*>(s_en + RC_SB * 32);  // [RC_SB][32] inclusive prefix of the words' survivor counts
    uint32_t* s_m = s_incl + RC_SB * 32;                                // [RC_SB][32] the bitmap words
    uint32_t* s_T = s_m + RC_SB * 32;                                   // [RC_SB] survivors per block
    uint32_t* s_sl = s_T + RC_SB + 2;                                   // [P] SINK = 2: the query's slot_local row (word -> probe rank)
    const uint32_t lt_mask = (1u << lane) - 1u;
    const long long dbg_t0 = a.dbg ? clock64() : 0ll;

    auto window = [&](uint32_t& wlo, uint32_t& whi) {  // word window of this round from the bounds (when no precomputed pair exists)
        const uint32_t wb = a.q_wbase[q], wend = a.q_wbase[q + 1];
        const int p0 = (int)a.q_p0[q];
        auto word_at = [&](int pe, int ch) -> uint32_t {
            const int p = pe + p0;
            if (p >= a.P) return wend;
            const uint32_t s0 = wb + a.slot_local[(size_t)q * a.P + p];
            const uint32_t s1 = (p + 1 < a.P) ? wb + a.slot_local[(size_t)q * a.P + p + 1] : wend;
            return min(s0 + (uint32_t)ch * (SCAN_THREADS / 32), s1);
        };
        wlo = word_at(p_lo, ch_lo); whi = word_at(p_hi, ch_hi);
    };
    volatile uint32_t* s_win = reinterpret_cast<volatile uint32_t*>(rc_smem_raw + 516);  // [3] first word, end word, (SINK = 2) the query's slot base
    if constexpr (SINK != 0) {  // most queries have no candidates in this shard's window: find out before anything else is loaded
        if (threadIdx.x == 0) {
            uint32_t wlo, whi;
            window(wlo, whi);
            if constexpr (SINK == 2) {  // nothing survived the frozen threshold, or the region overflowed (flagged by r2_offsets_kernel)
                const uint32_t c2 = a.r2_cnt[q];
                if (c2 == 0u || c2 == 0xffffffffu) whi = wlo;
            }
            s_win[0] = wlo; s_win[1] = whi;
            if constexpr (SINK == 2) s_win[2] = a.q_wbase[q];
        }
        __syncthreads();
        if (s_win[0] >= s_win[1]) return;
    }
    if (threadIdx.x < 48) mbar_init(&bar_full[threadIdx.x], 1);
    if (threadIdx.x == 0) *thr_s = (first && SINK != 2) ? 3.402823466e+38f : a.thr[q];  // SINK = 2: the frozen threshold of round 1
    if constexpr (SINK == 2) {
        for (int p = threadIdx.x; p < a.P; p += blockDim.x) s_sl[p] = a.slot_local[(size_t)q * a.P + p];
    }
    {
        const float4* src = reinterpret_cast<const float4*>(a.qpad + (size_t)q * D);
        float4* dst = reinterpret_cast<float4*>(qv);
        for (int d = threadIdx.x; d < D / 4; d += blockDim.x) dst[d] = __ldg(&src[d]);
    }
    __syncthreads();

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------- producer
        uint32_t wlo, whi;
        if constexpr (SINK != 0) {
            wlo = s_win[0]; whi = s_win[1];
        } else if (a.win) {
            const uint2 ww = a.win[q];
            wlo = ww.x; whi = ww.y;
        } else {
            window(wlo, whi);
        }
        uint32_t w = 0;
        int fill = 0;
        uint32_t dbg_blocked = 0;
        auto open_wave = [&]() {  // the wave's row buffer and record slot must have been drained
            const long long t0 = a.dbg ? clock64() : 0ll;
            if (w >= NS) mbar_wait<32>(&bar_rfree[w % NS], ((w / NS) + 1) & 1u);
            if (w >= NM) mbar_wait<32>(&bar_mfree[w % NM], ((w / NM) + 1) & 1u);
            if (a.dbg) dbg_blocked += (uint32_t)(clock64() - t0);
        };
        auto close_wave = [&](uint32_t last) {
            if (fill == 0) open_wave();
            __threadfence_block();
            __syncwarp();
            if (lane == 0) {
                wn[w % NM] = (uint32_t)fill;
                wlast[w % NM] = last;
                mbar_arrive(&bar_full[w % NS]);
            }
            w++;
            fill = 0;
        };
        auto enqueue = [&](uint32_t pm, float rough, uint32_t j, uint32_t word) {
            while (pm) {
                if (fill == 0) open_wave();
                const uint32_t s = w % NS, m = w % NM;
                const int space = R - fill;
                const int rank = __popc(pm & lt_mask);
                const bool take = ((pm >> lane) & 1u) && rank < space;
                const uint32_t took = __ballot_sync(FULL, take);
                if (lane == __ffs(took) - 1) mbar_expect_tx(&bar_full[s], rowbytes * (uint32_t)__popc(took));
                __syncwarp();
                if (take) {
                    const int slot = fill + rank;
                    qr[m * RW + slot] = rough;
                    qj[m * RW + slot] = j;
                    if constexpr (SINK == 2) sqw[m * RW + slot] = word;
                    tma_bulk_g2s(rows + ((size_t)s * R + slot) * pitch, a.base + (size_t)j * D, rowbytes, &bar_full[s]);
                }
                pm &= ~took;
                fill += __popc(took);
                if (fill == R) close_wave(0u);
            }
        };
        auto prefix = [&](uint32_t m) -> uint32_t {  // inclusive prefix sum of the words' survivor counts
            uint32_t x = __popc(m);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, x, o);
                if (lane >= o) x += t;
            }
            return x;
        };
        int pos_last = 0;  // word (inside its block) of the survivor the last load_chunk gave this lane
        auto load_chunk = [&](uint32_t w0, uint32_t incl, uint32_t exc, uint32_t T, uint32_t e0) -> float2 {  // survivor e0 + lane of the block
            const uint32_t e = e0 + lane;
            int pos = 0;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const uint32_t pv = __shfl_sync(FULL, incl, pos + s - 1);
                if (pv <= e) pos += s;
            }
            pos = min(pos, 31);
            const uint32_t src_excl = __shfl_sync(FULL, exc, pos);
            float2 en = make_float2(3.402823466e+38f, 0.0f);
            if (e < T) en = a.entries[(size_t)(w0 + pos) * 32 + (e - src_excl)];
            pos_last = pos;
            return en;
        };
        // Super-blocks of SB x 32 words.  The stream is a chain of dependent DRAM accesses (bitmap word -> entries -> row), about a
        // microsecond each under load, so it runs two super-blocks ahead: bitmap words of super-block i+2 and the first 32 entries
        // of every block of super-block i+1 are in flight while super-block i is filtered and queued (from a shared-memory stash,
        // so that the processing loop stays rolled).
        constexpr int SB = RC_SB;
        uint32_t mB[SB], mC[SB], inclB[SB];
        float2 enB[SB];
        auto load_bm = [&](uint32_t wbase, uint32_t (&m)[SB]) {
#pragma unroll
            for (int u = 0; u < SB; u++) {
                const uint32_t idx = wbase + u * 32 + lane;
                m[u] = idx < whi ? a.bitmap[idx] : 0u;
            }
        };
        auto load_entries = [&](uint32_t wbase) {  // needs mB
#pragma unroll
            for (int u = 0; u < SB; u++) inclB[u] = prefix(mB[u]);
#pragma unroll
            for (int u = 0; u < SB; u++)
                enB[u] = load_chunk(wbase + u * 32, inclB[u], inclB[u] - __popc(mB[u]), __shfl_sync(FULL, inclB[u], 31), 0u);
        };
        load_bm(wlo, mB);
        load_bm(wlo + 32 * SB, mC);
        load_entries(wlo);
        for (uint32_t w0 = wlo; w0 < whi; w0 += 32 * SB) {
#pragma unroll
            for (int u = 0; u < SB; u++) {
                s_en[u * 32 + lane] = enB[u];
                s_incl[u * 32 + lane] = inclB[u];
                s_m[u * 32 + lane] = mB[u];
                if (lane == 31) s_T[u] = inclB[u];
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < SB; u++) mB[u] = mC[u];
            load_bm(w0 + 2 * 32 * SB, mC);
            load_entries(w0 + 32 * SB);
#pragma unroll 1
            for (int u = 0; u < SB; u++) {
                const uint32_t T = s_T[u];
                if (T == 0) continue;  // uniform
                {
                    const float2 en = s_en[u * 32 + lane];
                    uint32_t word = 0;
                    if constexpr (SINK == 2) {  // the word of survivor `lane` of the block: first word whose inclusive prefix exceeds it
                        const uint32_t inc = s_incl[u * 32 + lane];
                        int pos = 0;
#pragma unroll
                        for (int sft = 16; sft > 0; sft >>= 1) {
                            const uint32_t pv = __shfl_sync(FULL, inc, pos + sft - 1);
                            if (pv <= (uint32_t)lane) pos += sft;
                        }
                        word = w0 + u * 32 + min(pos, 31);
                    }
                    const uint32_t pm = __ballot_sync(FULL, (uint32_t)lane < T && en.x < *thr_s);
                    if (pm) enqueue(pm, en.x, __float_as_uint(en.y), word);
                }
                if (T > 32) {  // dense block (loose threshold, e.g. the first probed cluster): further chunks, three at a time in flight
                    const uint32_t wblk = w0 + u * 32;
                    const uint32_t inc = s_incl[u * 32 + lane], exc = inc - __popc(s_m[u * 32 + lane]);
                    for (uint32_t e0 = 32; e0 < T; e0 += 96) {
                        float2 en[3];
                        uint32_t wd[3];
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            en[c] = load_chunk(wblk, inc, exc, T, e0 + 32 * c);
                            wd[c] = wblk + (uint32_t)pos_last;
                        }
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const uint32_t pm = __ballot_sync(FULL, e0 + 32 * c + lane < T && en[c].x < *thr_s);
                            if (pm) enqueue(pm, en[c].x, __float_as_uint(en[c].y), wd[c]);
                        }
                    }
                }
            }
            __syncwarp();
        }
        close_wave(1u);
        for (int i = 1; i < ncw; i++) close_wave(2u);  // one terminator for each of the other compute warps (waves go round-robin)
        if (a.dbg && lane == 0) {
            uint32_t* o = a.dbg + ((size_t)q * 2 + (first ? 0 : 1)) * 8;
            o[3] = dbg_blocked; o[7] = (uint32_t)(clock64() - dbg_t0);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------------- replay
        int cnt = first ? 0 : (int)a.heap_cnt[q];
        float thr = *thr_s;
        float hmax = 3.402823466e+38f;
        float recent = first ? -3.402823466e+38f : a.h_recent[q];
        uint32_t wcount = first ? 0u : a.h_wcount[q];
        int maxpos = 0;
        for (int s = lane; s < cnt; s += 32) {
            hd[s] = a.heap_dist[(size_t)q * k + s];
            hid[s] = a.heap_ids[(size_t)q * k + s];
        }
        __syncwarp();
        if (cnt == k) {
            heap_recompute_max(hd, k, lane, maxpos, hmax);
            if constexpr (!HEUR) thr = hmax;
        }
        uint32_t precise = 0, computed = 0, waves = 0, dbg_rwait = 0;
        // record sink (SINK = 1): the candidates the reference computes, in order, into this shard's region of the home rank's inbox
        const int home = SINK ? q / a.nq_local : 0, ql = SINK ? q - home * a.nq_local : 0;
        SurvRec* rdst = nullptr;
        uint32_t nrec = 0, prank = 0;
        if constexpr (SINK == 1) {
            rdst = reinterpret_cast<SurvRec*>(a.peers[home] + a.off_r1rec) + ((size_t)a.rank * a.nq_local + ql) * a.r1cap;
            prank = a.q_p0[q] + (uint32_t)p_lo;
        }
        // SINK = 2 (distributed frozen round, source side): this shard replays ITS candidates of the query in visit order against a
        // LOCAL threshold -- min(frozen threshold, k-th smallest max(exact, rough) among the shard's own computed candidates) -- and
        // ships every candidate it computes.  That threshold never drops below the one the reference holds at the same candidate:
        // k local keys below the reference's threshold would be k candidates the reference has computed as well (rough <= key <
        // its threshold, which only falls) with exact distances below its threshold, i.e. below its own k-th smallest.  So the
        // shipped set is a superset of what the reference reranks on this shard, far smaller than the frozen round's survivors.
        const float thr_cap = SINK == 2 ? thr : 3.402823466e+38f;
        if constexpr (SINK == 2)
            rdst = reinterpret_cast<SurvRec*>(a.peers[home] + a.off_r2rec) + (size_t)a.rank * a.cap2 + a.r2_off[q];
        for (uint32_t w = 0;; w++) {
            const uint32_t m = w % NM;
            const long long tw0 = a.dbg ? clock64() : 0ll;
            mbar_wait(&bar_exact[m], (w / NM) & 1u);
            if (a.dbg) dbg_rwait += (uint32_t)(clock64() - tw0);
            const int n = (int)wn[m];
            const uint32_t last = wlast[m];
            const bool mine = lane < n;
            const float rough = mine ? qr[m * RW + lane] : 0.0f;
            const uint32_t j = mine ? qj[m * RW + lane] : 0u;
            const float exact = mine ? ex[m * RW + lane] : 0.0f;
            uint32_t mid = 0;
            if constexpr (SINK != 0) mid = mine ? smid[m * RW + lane] : 0u;
            if constexpr (SINK == 2) prank = mine ? sqw[m * RW + lane] : 0u;
            // the key the heap keeps: the exact distance; SINK = 2: max(exact, rough) (a NaN distance is never kept, as in the reference)
            const float key = SINK == 2 ? (exact == exact ? fmaxf(exact, rough) : 3.402823466e+38f) : exact;
            computed += (uint32_t)n;
            waves += n ? 1u : 0u;
            // in-order replay (rerank.rs:83-101).  The threshold only moves when a candidate is ACCEPTED (rough < thr and
            // exact < thr), so the stretch up to the next accepted candidate is evaluated in one step: every lane of the
            // stretch with rough < thr is a candidate the reference computes an exact distance for.
            uint32_t rem = __ballot_sync(FULL, mine && rough < thr);
            while (rem) {
                const bool pass = ((rem >> lane) & 1u) && rough < thr;
                const uint32_t pm = __ballot_sync(FULL, pass);
                const uint32_t am = __ballot_sync(FULL, pass && key < thr);
                const int t = __ffs(am) - 1;                           // the next accepted candidate (-1: none)
                const uint32_t upto = am ? ((2u << t) - 1u) : FULL;    // lanes 0..t
                const uint32_t cm = pm & upto;
                precise += __popc(cm);
                if constexpr (SINK != 0) {
                    const uint32_t pos = nrec + __popc(cm & lt_mask);
                    if (((cm >> lane) & 1u) && (SINK == 2 || pos < (uint32_t)a.r1cap)) {  // (SINK = 2: at most the round's survivors, which the region holds)
                        SurvRec rec;
                        rec.rough = rough; rec.exact = exact; rec.id = mid; rec.p = prank;
                        rdst[pos] = rec;
                    }
                    nrec += __popc(cm);
                }
                if (!am) break;
                rem &= ~upto;
                const float exa = __shfl_sync(FULL, key, t);
                if (!HEUR || cnt < k || exa < hmax) {
                    const int slot = cnt < k ? cnt : maxpos;
                    if (lane == t) { hd[slot] = exa; hid[slot] = SINK == 1 ? mid : j; }  // SINK = 0: the POSITION; map_ids at finalize
                    if (cnt < k) cnt++;
                    __syncwarp();
                    if (cnt == k) {
                        heap_recompute_max(hd, k, lane, maxpos, hmax);
                        if constexpr (!HEUR) thr = SINK == 2 ? fminf(thr_cap, hmax) : hmax;  // rerank.rs:98-100
                    }
                }
                if constexpr (HEUR) {  // rerank.rs:155-162
                    wcount++;
                    recent = fmaxf(recent, exa);
                    if (wcount >= 12u) {
                        thr = recent;
                        wcount = 0;
                        recent = -3.402823466e+38f;
                    }
                }
                if (lane == 0) *thr_s = thr;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_mfree[m]);
            if (last) break;
        }
        if constexpr (SINK == 1) {  // the owner of the window tells the home rank how many records to replay
            if (lane == 0) reinterpret_cast<uint32_t*>(a.peers[home] + a.off_r1cnt)[(size_t)a.rank * a.nq_local + ql] = min(nrec, (uint32_t)a.r1cap);
        }
        if constexpr (SINK == 2) {  // (offset, count) of the query's run: r2_offsets_kernel wrote the survivor count, this is what was shipped
            if (lane == 0) reinterpret_cast<uint2*>(a.peers[home] + a.off_r2tab)[(size_t)a.rank * a.nq_local + ql] = make_uint2(a.r2_off[q], nrec);
        }
        if (a.dbg && lane == 0) {
            uint32_t* o = a.dbg + ((size_t)q * 2 + (first ? 0 : 1)) * 8;
            o[0] = (uint32_t)(clock64() - dbg_t0); o[1] = waves; o[2] = computed; o[6] = dbg_rwait;
        }
        if (lane == 0) {
            a.q_precise[q] = (first ? 0u : a.q_precise[q]) + precise;
            atomicAdd(&a.counters[1], (unsigned long long)computed);
            if (SINK != 2) atomicAdd(&a.counters[2], (unsigned long long)precise);
        }
        if (SINK == 2) {
            // (the local heap dies with the CTA: the home rank replays the records against the real threshold)
        } else if (!finalize) {
            for (int s = lane; s < cnt; s += 32) {
                a.heap_dist[(size_t)q * k + s] = hd[s];
                a.heap_ids[(size_t)q * k + s] = hid[s];
            }
            if (lane == 0) { a.heap_cnt[q] = (uint32_t)cnt; a.thr[q] = thr; a.h_recent[q] = recent; a.h_wcount[q] = wcount; }
        } else {
            // ascending by (distance, id): rank by counting; the original ids (rerank.rs:94) are looked up here
            __syncwarp();
            if constexpr (SINK == 0) {
                for (int s = lane; s < cnt; s += 32) hid[s] = a.map_ids[hid[s]];
                __syncwarp();
            }
            for (int s = lane; s < k; s += 32) {
                if (s < cnt) {
                    const uint32_t ks = okey(hd[s]), is = hid[s];
                    int rank = 0;
                    for (int t = 0; t < cnt; t++) {
                        const uint32_t kt = okey(hd[t]), itt = hid[t];
                        rank += (kt < ks) || (kt == ks && (itt < is || (itt == is && t < s)));
                    }
                    a.out_dist[(size_t)q * k + rank] = hd[s];
                    a.out_ids[(size_t)q * k + rank] = is;
                } else {
                    a.out_dist[(size_t)q * k + s] = __int_as_float(0x7f800000);
                    a.out_ids[(size_t)q * k + s] = 0xffffffffu;
                }
            }
            if (lane == 0) a.out_count[q] = (uint32_t)cnt;
        }
    } else {
        // ------------------------------------------------------------------------------------------- compute
        // waves go round-robin over the compute warps: a wave (<= 4 NC rows) is one warp's job, several waves are computed at once
        const int cw = warp - 2, grp = lane >> 3, l8 = lane & 7;
        uint32_t dbg_cwait = 0, dbg_cl2 = 0;
        for (uint32_t w = (uint32_t)cw;; w += (uint32_t)ncw) {
            const uint32_t s = w % NS, m = w % NM;
            const long long tc0 = a.dbg ? clock64() : 0ll;
            mbar_wait(&bar_full[s], (w / NS) & 1u);
            const long long tc1 = a.dbg ? clock64() : 0ll;
            dbg_cwait += (uint32_t)(tc1 - tc0);
            const int n = (int)wn[m];
            const uint32_t last = wlast[m];
            const float thr = *thr_s;  // stale at worst = larger: whatever the replay will test has been computed
            const float* rw = rows + (size_t)s * R * pitch;
            int c[NC];
            bool need = false;
#pragma unroll
            for (int i = 0; i < NC; i++) {
                c[i] = i * 4 + grp;
                need = need || (c[i] < n && qr[m * RW + c[i]] < thr);
            }
            if (__any_sync(FULL, need)) {
                uint32_t midv[NC];
                if constexpr (SINK != 0) {  // in flight while the distances are computed
#pragma unroll
                    for (int i = 0; i < NC; i++) midv[i] = (l8 == 0 && c[i] < n) ? a.map_ids[qj[m * RW + c[i]]] : 0u;
                }
                if constexpr (SINK == 2) {  // probe rank of the slot holding the candidate's word: the last p with slot_local[p] <= word - base
#pragma unroll
                    for (int i = 0; i < NC; i++)
                        if (l8 == 0 && c[i] < n) {
                            const uint32_t rel = sqw[m * RW + c[i]] - s_win[2];
                            int lo = 0, hi = a.P;
                            while (hi - lo > 1) {
                                const int md = (lo + hi) >> 1;
                                if (s_sl[md] <= rel) lo = md; else hi = md;
                            }
                            sqw[m * RW + c[i]] = (uint32_t)lo;
                        }
                }
                float d2[NC];
                if constexpr (NC == 2) {
                    const float* const rp[2] = {rw + (size_t)min(c[0], R - 1) * pitch, rw + (size_t)min(c[1], R - 1) * pitch};
                    l2_oct<2>(rp, qv, D, l8, d2);
                } else {
                    const float* const rp[1] = {rw + (size_t)min(c[0], R - 1) * pitch};
                    l2_oct<1>(rp, qv, D, l8, d2);
                }
#pragma unroll
                for (int i = 0; i < NC; i++)
                    if (l8 == 0 && c[i] < n) {
                        ex[m * RW + c[i]] = d2[i];
                        if constexpr (SINK != 0) smid[m * RW + c[i]] = midv[i];
                    }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the rows of this buffer are rewritten by later bulk copies
            __threadfence_block();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&bar_rfree[s]); mbar_arrive(&bar_exact[m]); }
            if (a.dbg) dbg_cl2 += (uint32_t)(clock64() - tc1);
            if (last) break;
        }
        if (a.dbg && cw == 0 && lane == 0) {
            uint32_t* o = a.dbg + ((size_t)q * 2 + (first ? 0 : 1)) * 8;
            o[4] = dbg_cwait; o[5] = dbg_cl2;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K6 (after the NCCL all-gather of per-shard results): k smallest of n_lists x topk candidates per query.
__global__ void merge_topk_kernel(const float* __restrict__ dist, const uint32_t* __restrict__ ids, int n_lists, size_t nq, int topk,
                                  float* __restrict__ out_dist, uint32_t* __restrict__ out_ids, uint32_t* __restrict__ out_count) {
    const int lane = threadIdx.x & 31;
    const size_t q = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int total = n_lists * topk;
    uint32_t cnt = 0;
    for (int s = lane; s < total; s += 32) {
        const int l = s / topk, i = s % topk;
        const size_t src = ((size_t)l * nq + q) * topk + i;
        const uint32_t is = ids[src];
        if (is == 0xffffffffu) continue;  // unused tail of a shard's list
        const uint32_t ks = okey(dist[src]);
        int rank = 0;
        for (int t = 0; t < total; t++) {
            const size_t o = ((size_t)(t / topk) * nq + q) * topk + (t % topk);
            const uint32_t it = ids[o];
            if (it == 0xffffffffu) continue;
            const uint32_t kt = okey(dist[o]);
            rank += (kt < ks) || (kt == ks && (it < is || (it == is && t < s)));
        }
        if (rank < topk) {
            out_dist[q * topk + rank] = dist[src];
            out_ids[q * topk + rank] = is;
        }
        cnt++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    const uint32_t have = min(cnt, (uint32_t)topk);
    for (int s = have + lane; s < topk; s += 32) {
        out_dist[q * topk + s] = __int_as_float(0x7f800000);
        out_ids[q * topk + s] = 0xffffffffu;
    }
    if (lane == 0) out_count[q] = have;
}

// =========================================================================================================
// Distributed pipeline: the index is sharded by cluster range, every rank is the HOME of a
// contiguous slice of the query batch (front end + final replay) and a SOURCE of survivor records for all queries.

// Layout of one rank's all-gather chunk (4-byte words, every section padded to 16 bytes):
//   [q: nq_l x len][y: nq_l x D][probe ids: nq_l x P][probe dist: nq_l x P][p0: nq_l]
struct DistChunk {
    size_t o_q, o_y, o_ids, o_dist, o_p0, words;
    // the same products as TWO chunks, so that the (big) q/y part can be all-gathered while the centroid scan still runs:
    // part A = [q | y] (words_a), part B = [ids | dist | p0] (words_b); a_* / b_* are offsets inside the respective part
    size_t a_q, a_y, words_a, b_ids, b_dist, b_p0, words_b;
};
__host__ __device__ inline DistChunk dist_chunk_layout(size_t nq_l, size_t len, size_t D, size_t P) {
    auto up4 = [](size_t x) { return (x + 3) & ~(size_t)3; };
    DistChunk c;
    c.o_q = 0;
    c.o_y = up4(c.o_q + nq_l * len);
    c.o_ids = up4(c.o_y + nq_l * D);
    c.o_dist = up4(c.o_ids + nq_l * P);
    c.o_p0 = up4(c.o_dist + nq_l * P);
    c.words = up4(c.o_p0 + nq_l);
    c.a_q = 0;
    c.a_y = c.o_y;
    c.words_a = c.o_ids;
    c.b_ids = 0;
    c.b_dist = c.o_dist - c.o_ids;
    c.b_p0 = c.o_p0 - c.o_ids;
    c.words_b = c.words - c.o_ids;
    return c;
}

// After the all-gather of the per-rank front-end products: `gathered` = world chunks -> flat per-query arrays over the
// whole batch (global query gq = rank * nq_l + i), queries zero-padded to D.  One warp per query, 128-bit copies.
__global__ void dist_unpack_kernel(const uint32_t* __restrict__ gathered_a, size_t stride_a, const uint32_t* __restrict__ gathered_b,
                                   size_t stride_b, int world, int nq_l, int len, int D, int P, float* __restrict__ qpad,
                                   float* __restrict__ y, uint32_t* __restrict__ probe_ids, float* __restrict__ probe_dist,
                                   uint32_t* __restrict__ q_p0) {
    // gathered_a / stride_a: rank r's [q | y] part starts at gathered_a + r * stride_a; gathered_b likewise for [ids | dist | p0]
    // (one buffer with the combined chunk layout, or the two buffers of the split all-gather)
    const int lane = threadIdx.x & 31;
    const size_t gq = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gq >= (size_t)world * nq_l) return;
    const DistChunk L = dist_chunk_layout(nq_l, len, D, P);
    const int r = (int)(gq / nq_l), ql = (int)(gq % nq_l);
    const uint32_t* cb = gathered_b + (size_t)r * stride_b;
    if (qpad) {  // (NULL: the rows were pushed straight into their flat arrays, only the probe lists travel by collective)
    const uint32_t* ca = gathered_a + (size_t)r * stride_a;
    {
        const uint4* src = reinterpret_cast<const uint4*>(ca + L.a_y + (size_t)ql * D);
        uint4* dst = reinterpret_cast<uint4*>(y + gq * D);
        for (int i = lane; i < D / 4; i += 32) dst[i] = src[i];
    }
    if ((len & 3) == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(ca + L.a_q + (size_t)ql * len);
        uint4* dst = reinterpret_cast<uint4*>(qpad + gq * D);
        for (int i = lane; i < D / 4; i += 32) dst[i] = i < len / 4 ? src[i] : make_uint4(0u, 0u, 0u, 0u);
    } else {
        const uint32_t* src = ca + L.a_q + (size_t)ql * len;
        for (int i = lane; i < D; i += 32) qpad[gq * D + i] = i < len ? __uint_as_float(src[i]) : 0.0f;
    }
    }
    for (int i = lane; i < P; i += 32) {
        probe_ids[gq * P + i] = cb[L.b_ids + (size_t)ql * P + i];
        probe_dist[gq * P + i] = __uint_as_float(cb[L.b_dist + (size_t)ql * P + i]);
    }
    if (lane == 0) q_p0[gq] = cb[L.b_p0 + ql];
}

// Survivor-slot layout of one shard for probe lists that were selected elsewhere: per (query, rank) the exclusive prefix
// of 32-vector words of the probed clusters THIS shard holds, and the per-query totals (what select_probe_kernel emits
// when it runs on the shard itself).
__global__ void __launch_bounds__(SEL_THREADS) slot_layout_kernel(const uint32_t* __restrict__ probe_ids, const uint32_t* __restrict__ offsets,
                                                                  int P, uint32_t* __restrict__ slot_local, uint32_t* __restrict__ q_words,
                                                                  uint32_t* __restrict__ q_pairs) {
    extern __shared__ uint32_t sl_words[];  // P
    __shared__ uint32_t warp_tot[SEL_THREADS / 32 + 1];
    __shared__ uint32_t s_pairs;
    const int tid = threadIdx.x;
    const size_t q = blockIdx.x;
    if (tid == 0) s_pairs = 0;
    __syncthreads();
    uint32_t pairs = 0;
    for (int p = tid; p < P; p += SEL_THREADS) {
        const uint32_t c = probe_ids[q * P + p], n_c = offsets[c + 1] - offsets[c];
        sl_words[p] = (n_c + 31u) >> 5;
        pairs += n_c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(FULL, pairs, o);
    if ((tid & 31) == 0 && pairs) atomicAdd(&s_pairs, pairs);
    __syncthreads();
    const uint32_t total = block_exclusive_scan<SEL_THREADS>(sl_words, P, warp_tot);
    for (int p = tid; p < P; p += SEL_THREADS) slot_local[q * P + p] = sl_words[p];
    if (tid == 0) { q_words[q] = total; q_pairs[q] = s_pairs; }
}

// Records this shard will ship per query in the frozen round: the survivors the scan left in the round's word window.
__global__ void r2_count_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ q_wbase, const uint32_t* __restrict__ slot_local,
                                const uint32_t* __restrict__ q_p0, int nq, int P, int p_lo, int ch_lo, uint32_t* __restrict__ r2_cnt) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const uint32_t wb = q_wbase[q], wend = q_wbase[q + 1];
    const int p = p_lo + (int)q_p0[q];
    uint32_t wlo = wend;
    if (p < P) {
        const uint32_t s0 = wb + slot_local[(size_t)q * P + p];
        const uint32_t s1 = (p + 1 < P) ? wb + slot_local[(size_t)q * P + p + 1] : wend;
        wlo = min(s0 + (uint32_t)ch_lo * (SCAN_THREADS / 32), s1);
    }
    uint32_t c = 0;
    for (uint32_t w = wlo + lane; w < wend; w += 32) c += __popc(bitmap[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if (lane == 0) r2_cnt[q] = c;
}

// One block per home rank: exclusive scan of the counts of the queries homed there = where each query's records start
// inside this shard's region of that rank's inbox; the (offset, count) table is written into the inbox as well.  A region
// that would overflow marks the affected queries (count 0xffffffff) and raises the status word: the step is then repeated
// with larger regions by the host, never silently truncated.
__global__ void __launch_bounds__(1024) r2_offsets_kernel(uint32_t* __restrict__ r2_cnt, uint32_t* __restrict__ r2_off, int nq_l, uint32_t cap2,
                                                          unsigned char* const* __restrict__ peers, size_t off_r2tab, int rank,
                                                          uint32_t* __restrict__ status, uint32_t* __restrict__ home_tot) {
    __shared__ uint32_t wtot[33];
    __shared__ uint32_t s_eff;
    if (threadIdx.x == 0) s_eff = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int home = blockIdx.x;
    uint32_t* cnt = r2_cnt + (size_t)home * nq_l;
    uint32_t* off = r2_off + (size_t)home * nq_l;
    uint2* tab = reinterpret_cast<uint2*>(peers[home] + off_r2tab) + (size_t)rank * nq_l;
    const int per = (nq_l + 1023) / 1024;
    const int lo = min(nq_l, tid * per), hi = min(nq_l, lo + per);
    uint32_t s = 0;
    for (int i = lo; i < hi; i++) s += cnt[i];
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wtot[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(FULL, wi, o);
            if (lane >= o) wi += v;
        }
        wtot[lane] = wi - w;
        if (lane == 31) wtot[32] = wi;
    }
    __syncthreads();
    uint32_t run = wtot[warp] + inc - s, eff = 0;
    for (int i = lo; i < hi; i++) {
        const uint32_t c = cnt[i];
        off[i] = run;
        if ((unsigned long long)run + c > cap2) {  // the running offset only grows: every later query with records is flagged too,
            cnt[i] = 0xffffffffu;                  // so the records that ARE shipped stay contiguous from the region start
            tab[i] = make_uint2(run, 0xffffffffu);
            atomicOr(status, 1u);
        } else {
            tab[i] = make_uint2(run, c);
            eff += c;
        }
        run += c;
    }
    if (eff) atomicAdd(&s_eff, eff);
    __syncthreads();
    if (tid == 0) home_tot[home] = s_eff;  // records this shard ships to `home` = length of its run in the flat candidate list
}

// Flat candidate list of the frozen round: (rough, row, query, probe rank) of every survivor, home-major, then query-major,
// then visit order -- the order the records will have in the home inboxes.  One warp per query streams its survivor words.
struct __align__(16) Cand {
    float rough;
    uint32_t j, q, p;
};

__global__ void r2_compact_kernel(const uint32_t* __restrict__ bitmap, const float2* __restrict__ entries, const uint32_t* __restrict__ q_wbase,
                                  const uint32_t* __restrict__ slot_local, const uint32_t* __restrict__ q_p0, const uint32_t* __restrict__ r2_cnt,
                                  const uint32_t* __restrict__ r2_off, const uint32_t* __restrict__ home_tot, int nq, int nq_l, int P,
                                  int p_lo, int ch_lo, Cand* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const uint32_t c2 = r2_cnt[q];
    if (c2 == 0u || c2 == 0xffffffffu) return;  // nothing to ship / flagged overflow
    const int home = q / nq_l;
    uint32_t base = r2_off[q];
    for (int h = 0; h < home; h++) base += home_tot[h];
    const uint32_t wb = q_wbase[q], wend = q_wbase[q + 1];
    const uint32_t* sl = slot_local + (size_t)q * P;
    const int p = p_lo + (int)q_p0[q];
    uint32_t wlo = wend;
    if (p < P) {
        const uint32_t s0 = wb + sl[p], s1 = (p + 1 < P) ? wb + sl[p + 1] : wend;
        wlo = min(s0 + (uint32_t)ch_lo * (SCAN_THREADS / 32), s1);
    }
    uint32_t run = 0;
    for (uint32_t w0 = wlo; w0 < wend; w0 += 32) {
        const uint32_t m = (w0 + lane < wend) ? bitmap[w0 + lane] : 0u;
        const uint32_t pc = __popc(m);
        uint32_t incl = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t T = __shfl_sync(FULL, incl, 31);
        for (uint32_t e0 = 0; e0 < T; e0 += 32) {
            const uint32_t e = e0 + lane;
            int pos = 0;  // first word whose inclusive prefix exceeds e
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const uint32_t pv = __shfl_sync(FULL, incl, pos + s - 1);
                if (pv <= e) pos += s;
            }
            pos = min(pos, 31);
            const uint32_t src_excl = __shfl_sync(FULL, incl - pc, pos);
            if (e < T) {
                const float2 en = entries[(size_t)(w0 + pos) * 32 + (e - src_excl)];
                const uint32_t rel = w0 + pos - wb;  // probe rank of the slot holding this word: the last p with slot_local[p] <= rel
                int lo = 0, hi = P;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (sl[mid] <= rel) lo = mid; else hi = mid;
                }
                Cand c;
                c.rough = en.x; c.j = __float_as_uint(en.y); c.q = (uint32_t)q; c.p = (uint32_t)lo;
                out[(size_t)base + run + e] = c;
            }
        }
        run += T;
    }
}

// Exact squared L2 of every candidate of the flat list against its (unrotated, padded) query, in the order of
// simd::l2_squared_distance (src/simd.rs:14-73): 8 lanes = the 8 AVX lanes, lane v walks elements v, v+8, ... with one
// fused multiply-add each, then reduce_f32_256.  Load-balanced over candidates, not queries (survivor counts are skewed).
// The result is stored straight into the inbox of the query's home rank: peer memory over NVLink when that is another GPU.
__global__ void __launch_bounds__(256) r2_exact_kernel(const Cand* __restrict__ cand, const uint32_t* __restrict__ home_tot, int world,
                                                       const float* __restrict__ qpad, const float* __restrict__ base,
                                                       const uint32_t* __restrict__ map_ids, int D, int nq_l, int rank, uint32_t cap2,
                                                       unsigned char* const* __restrict__ peers, size_t off_r2rec,
                                                       unsigned long long* __restrict__ counters) {
    uint32_t total = 0;
    for (int h = 0; h < world; h++) total += home_tot[h];
    // TWO threads per candidate: thread h owns AVX lanes 4h..4h+3 of simd::l2_squared_distance as two packed f32x2 chains and reads
    // its half of every 8-element step with one 128-bit load (row and query), so a warp keeps 16 rows streaming.
    const int h = threadIdx.x & 1, sub = (threadIdx.x & 31) >> 1;
    const uint32_t groups = (gridDim.x * blockDim.x) >> 1;
    for (uint32_t gb = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 4; gb < total; gb += groups) {  // warp-uniform trip count
        const uint32_t g = gb + sub;
        const bool act = g < total;
        const Cand c = cand[act ? g : total - 1];
        const ulonglong2* rp = reinterpret_cast<const ulonglong2*>(base + (size_t)c.j * D) + h;
        const ulonglong2* qp = reinterpret_cast<const ulonglong2*>(qpad + (size_t)c.q * D) + h;
        f32x2 a0 = 0ull, a1 = 0ull;
#pragma unroll 8  // 16 independent 128-bit loads in flight per thread: the chain is a sequence of DRAM round trips otherwise
        for (int i = 0; i < D / 8; i++) {
            const ulonglong2 x = __ldg(&rp[2 * i]), y = __ldg(&qp[2 * i]);
            const f32x2 f0 = sub2(x.x, y.x), f1 = sub2(x.y, y.y);  // diff = row - q (src/simd.rs:34-37)
            a0 = fma2(f0, f0, a0);
            a1 = fma2(f1, f1, a1);
        }
        float p0, p1, p2, p3, o0, o1, o2, o3;  // reduce_f32_256: (s0+s4 + s1+s5) + (s2+s6 + s3+s7); both threads get the same bits
        unpack2(a0, p0, p1); unpack2(a1, p2, p3);
        unpack2(__shfl_xor_sync(FULL, a0, 1), o0, o1); unpack2(__shfl_xor_sync(FULL, a1, 1), o2, o3);
        const float acc = __fadd_rn(__fadd_rn(__fadd_rn(p0, o0), __fadd_rn(p1, o1)), __fadd_rn(__fadd_rn(p2, o2), __fadd_rn(p3, o3)));
        if (h == 0 && act) {
            const int home = (int)(c.q / (uint32_t)nq_l);
            uint32_t hb = 0;
            for (int hh = 0; hh < home; hh++) hb += home_tot[hh];
            SurvRec rec;
            rec.rough = c.rough; rec.exact = acc; rec.id = map_ids[c.j]; rec.p = c.p;
            reinterpret_cast<SurvRec*>(peers[home] + off_r2rec)[(size_t)rank * cap2 + (g - hb)] = rec;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counters[1], (unsigned long long)total);
}

// Final replay on the HOME rank: HeapReRanker::rank_batch (src/rerank.rs:81-106) over the union of the records every shard
// shipped, in the reference's visit order -- round-1 records first (they are a prefix of the order), then the frozen
// round's records merged by probe rank (a rank's cluster lives on exactly one shard, and a shard's records are already in
// its visit order).  The union is a superset of what the reference reranks (frozen threshold >= every later threshold),
// so the two strict tests below see exactly the candidates, thresholds and heap states of the sequential loop: ids,
// distances and the `precise` counter are the reference's.
struct HomeArgs {
    const unsigned char* inbox;   // this rank's inbox
    size_t off_r1cnt, off_r1rec, off_r2tab, off_r2rec;
    const uint32_t* probe_ids;    // whole batch, nq x P
    const uint32_t* q_p0;         // whole batch
    const uint32_t* goffsets;     // K+1, global rows
    const uint32_t* row_bounds;   // world+1, global rows of the shards
    float* out_dist;              // nq_l x topk
    uint32_t* out_ids;
    uint32_t* out_count;
    unsigned long long* counters; // [2] precise
    uint32_t* status;
    int world, rank, nq_l, P, topk, r1cap;
    uint32_t cap2;
};

__global__ void __launch_bounds__(128) home_replay_kernel(HomeArgs a) {
    extern __shared__ __align__(16) unsigned char hr_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int ql = blockIdx.x * wpb + warp;
    if (ql >= a.nq_l) return;
    const int k = a.topk;
    float* hd = reinterpret_cast<float*>(hr_smem) + (size_t)warp * 2 * k;
    uint32_t* hid = reinterpret_cast<uint32_t*>(hd + k);
    const size_t gq = (size_t)a.rank * a.nq_l + ql;
    const int p0 = (int)a.q_p0[gq];
    int cnt = 0, maxpos = 0;
    float thr = 3.402823466e+38f;
    uint32_t precise = 0;

    auto replay = [&](const SurvRec& rec, int n) {  // lanes 0..n-1 hold consecutive records
        // the threshold only decreases: a record that fails `rough < thr` now fails it later too, so only these are visited
        uint32_t rem = __ballot_sync(FULL, lane < n && rec.rough < thr);
        while (rem) {
            const int t = __ffs(rem) - 1;
            rem &= rem - 1;
            const float r = __shfl_sync(FULL, rec.rough, t), ex = __shfl_sync(FULL, rec.exact, t);
            const uint32_t id = __shfl_sync(FULL, rec.id, t);
            if (r < thr) {            // rerank.rs:84
                precise++;
                if (ex < thr) {       // rerank.rs:92
                    const int slot = cnt < k ? cnt : maxpos;
                    if (lane == 0) { hd[slot] = ex; hid[slot] = id; }
                    if (cnt < k) cnt++;
                    __syncwarp();
                    if (cnt == k) heap_recompute_max(hd, k, lane, maxpos, thr);  // rerank.rs:98-100
                }
            }
        }
    };
    auto owner_of = [&](uint32_t c) -> int {
        const uint32_t row = a.goffsets[c];
        int s = 0;
        while (s + 1 < a.world && row >= a.row_bounds[s + 1]) s++;
        return s;
    };

    bool bad = false;
    if (p0 < a.P) {
        // ---- round 1: written by the shard that owns the nearest non-empty cluster
        const uint32_t c0 = a.probe_ids[gq * a.P + p0];
        if (a.goffsets[c0 + 1] > a.goffsets[c0]) {
            const int o1 = owner_of(c0);
            const uint32_t n1 = reinterpret_cast<const uint32_t*>(a.inbox + a.off_r1cnt)[(size_t)o1 * a.nq_l + ql];
            const SurvRec* r1 = reinterpret_cast<const SurvRec*>(a.inbox + a.off_r1rec) + ((size_t)o1 * a.nq_l + ql) * a.r1cap;
            for (uint32_t i0 = 0; i0 < n1; i0 += 32) {
                SurvRec rec = {0.f, 0.f, 0u, 0u};
                if (i0 + lane < n1) rec = r1[i0 + lane];
                replay(rec, (int)min(32u, n1 - i0));
            }
        }
        // ---- frozen round: merge the shards' runs by probe rank
        uint32_t soff = 0, scnt = 0, scur = 0, head = 0xffffffffu;
        const SurvRec* r2 = reinterpret_cast<const SurvRec*>(a.inbox + a.off_r2rec);
        if (lane < a.world) {
            const uint2 t = reinterpret_cast<const uint2*>(a.inbox + a.off_r2tab)[(size_t)lane * a.nq_l + ql];
            soff = t.x; scnt = t.y;
            if (scnt == 0xffffffffu) { bad = true; scnt = 0; }
            if (scnt) head = r2[(size_t)lane * a.cap2 + soff].p;
        }
        bad = __any_sync(FULL, bad);
        while (!bad) {
            uint32_t pm = head;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pm = min(pm, __shfl_xor_sync(FULL, pm, o));
            if (pm == 0xffffffffu) break;
            const int s = __ffs(__ballot_sync(FULL, head == pm)) - 1;
            const uint32_t o_s = __shfl_sync(FULL, soff, s), c_s = __shfl_sync(FULL, scnt, s);
            uint32_t cur = __shfl_sync(FULL, scur, s);
            uint32_t next_head = 0xffffffffu;
            for (;;) {
                SurvRec rec = {0.f, 0.f, 0u, 0xffffffffu};
                if (cur + lane < c_s) rec = r2[(size_t)s * a.cap2 + o_s + cur + lane];
                const uint32_t same = __ballot_sync(FULL, rec.p == pm);
                const int n = same == FULL ? 32 : __ffs(~same) - 1;  // records are rank-ordered: a prefix
                replay(rec, n);
                cur += n;
                if (n < 32) {
                    next_head = __shfl_sync(FULL, rec.p, n);  // 0xffffffff past the end of the segment
                    break;
                }
                if (cur >= c_s) break;
            }
            if (lane == s) { scur = cur; head = next_head; }
        }
    }
    if (bad) {
        if (lane == 0) atomicOr(a.status, 2u);
        cnt = 0;
    }
    __syncwarp();
    for (int s = lane; s < k; s += 32) {
        if (s < cnt) {
            const uint32_t ks = okey(hd[s]), is = hid[s];
            int rank = 0;
            for (int t = 0; t < cnt; t++) {
                const uint32_t kt = okey(hd[t]), itt = hid[t];
                rank += (kt < ks) || (kt == ks && (itt < is || (itt == is && t < s)));
            }
            a.out_dist[(size_t)ql * k + rank] = hd[s];
            a.out_ids[(size_t)ql * k + rank] = is;
        } else {
            a.out_dist[(size_t)ql * k + s] = __int_as_float(0x7f800000);
            a.out_ids[(size_t)ql * k + s] = 0xffffffffu;
        }
    }
    if (lane == 0) {
        a.out_count[ql] = (uint32_t)cnt;
        atomicAdd(&a.counters[2], (unsigned long long)precise);
    }
}

// all-reduce(min) of the round-1 thresholds when every rank lives in ONE process (tests on a single GPU): plain device code.
__global__ void min_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = fminf(dst[i], src[i]);
}

__global__ void fill_f32_kernel(float* p, size_t n, float v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace rq
