// prefilter.cuh -- tensor-core prefilter for the centroid scan + probe selection (src/rabitq.rs:283-297).
//
// The reference computes all K exact fp32 distances ||c - y||^2 per query (simd::l2_squared_distance, src/simd.rs:14-73)
// and keeps the `probe` nearest.  The probe LIST must be exactly the reference's (the quantised query planes depend on it),
// but that only needs exact distances for the centroids that can be among the `probe` nearest.  So:
//   1. approximate keys  A[q][c] = ||c'||^2 - 2 <y', c'>  with y' = y - mu, c' = c - mu (mu = mean centroid: distances are
//      translation invariant and centring shrinks the norms the error bound scales with), the inner products on the tensor
//      cores in TF32 (operands rounded to TF32 once, fp32 accumulation);
//   2. a RIGOROUS bound  |A[q][c] + ||y'||^2 - e[q][c]| <= m[q][c] = alpha ||y'|| ||c'|| + 2^-21 (||y'|| + ||c'||)^2
//      + gamma (A[q][c] + ||y'||^2)  on the difference to the reference's fp32 value e:  alpha = 2^-8 for plain TF32 operands
//      (operand rounding 2 * 2^-11 and tensor-core accumulation <= D 2^-23 on each of the two sides, 1.6x slack) or
//      (D + 64) 2^-21 for the 3xTF32 split (hi*hi + hi*lo + lo*hi: only the accumulation term is left, 4x slack); the second
//      term covers the fp32 assembly of A; gamma = (D/8 + 8) 2^-23 the fp32 evaluation of the reference's own formula.  The
//      handle starts with plain TF32 and moves to 3xTF32 (then to the exact path) when a batch cannot be certified;
//   3. per query: a threshold tau with #{c : A_c <= tau} >= probe (256-bin histogram of the keys, verified by an exact
//      count), candidates = {c : A_c - m_c <= tau + max_c m_c} -- a superset of the reference's probe set, ties included --
//      EXACT distances for the candidates only, in the reference's AVX order, then the `probe` smallest by (distance, id).
// The output is bit-identical to centroid_dist_kernel + select_probe_kernel; a query whose candidate set overflows raises a
// flag and the batch is redone on that classic path.
#pragma once

#include "kernels.cuh"

namespace rq {

RQ_DEV float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---- index side: mu, c' = tf32(c - mu), ||c'||, ||c'||^2 ------------------------------------------------------------------
__global__ void centroid_mean_kernel(const float* __restrict__ cent, int K, int D, float* __restrict__ mu) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    double s = 0.0;
    for (int c = 0; c < K; c++) s += (double)cent[(size_t)c * D + d];
    mu[d] = (float)(s / (double)K);
}

__global__ void centroid_center_kernel(const float* __restrict__ cent, const float* __restrict__ mu, int K, int D, float* __restrict__ chat,
                                       float* __restrict__ chat_lo, float* __restrict__ cnorm, float* __restrict__ cnorm2,
                                       float* __restrict__ cnorm_max) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= K) return;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) {
        const float v = __fsub_rn(cent[(size_t)c * D + d], mu[d]);
        const float h = to_tf32(v);
        chat[(size_t)c * D + d] = h;
        chat_lo[(size_t)c * D + d] = to_tf32(__fsub_rn(v, h));  // 3xTF32 split: v = hi + lo up to 2^-22 |v|
        s += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) {
        const float n2 = (float)s, n = (float)sqrt(s) * 1.0000005f;  // norm rounded UP: it only ever enters upper bounds
        cnorm2[c] = n2;
        cnorm[c] = n;
        atomicMax(reinterpret_cast<int*>(cnorm_max), __float_as_int(n));  // non-negative floats order like ints
    }
}

// ---- query side: y^ = tf32(y - mu), ||y - mu|| -----------------------------------------------------------------------------
__global__ void query_center_kernel(const float* __restrict__ y, const float* __restrict__ mu, int nq, int D, float* __restrict__ yhat,
                                    float* __restrict__ yhat_lo, float* __restrict__ ynorm) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    float s = 0.0f;
    for (int d = lane; d < D; d += 32) {
        const float v = __fsub_rn(y[(size_t)q * D + d], mu[d]);
        const float h = to_tf32(v);
        yhat[(size_t)q * D + d] = h;
        if (yhat_lo) yhat_lo[(size_t)q * D + d] = to_tf32(__fsub_rn(v, h));
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) ynorm[q] = __fsqrt_ru(s) * 1.00001f;  // rounded up (32-term fp32 sum: relative error << 1e-5)
}

// ---- A[q][c] = ||c'||^2 - 2 <y^, c^> on the tensor cores (mma.sync m16n8k8 TF32, fp32 accumulate) ----------------------------
// CTA tile BM queries x BN centroids, K-chunks of 32 staged in shared memory with cp.async (2 stages); every warp owns a
// (BM/WM) x (BN/WN) sub-tile.  Both operands are d-contiguous ("K-major"); rows are padded to 36 floats so that the 8 x 4
// lanes of a fragment load hit 32 different banks.
constexpr int PF_BK = 32;
constexpr int PF_PITCH = PF_BK + 4;

// (cp_async16 / cp_async_commit / cp_async_wait: kernels.cuh)

template <int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32) approx_gemm_tf32_kernel(const float* __restrict__ yhat, const float* __restrict__ chat,
                                                                        const float* __restrict__ cnorm2, int nq, int K, int D,
                                                                        float* __restrict__ A, int accumulate) {
    constexpr int THREADS = WM * WN * 32;
    constexpr int TM = BM / WM / 16, TN = BN / WN / 8;  // mma tiles per warp
    extern __shared__ __align__(16) float pf_smem[];
    float* sA = pf_smem;                              // [2][BM][PITCH]
    float* sB = pf_smem + 2 * BM * PF_PITCH;          // [2][BN][PITCH]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
    float acc[TM][TN][4];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[i][j][e] = 0.0f;

    auto load_stage = [&](int stage, int d0) {
        float* a = sA + stage * BM * PF_PITCH;
        float* b = sB + stage * BN * PF_PITCH;
        for (int i = tid; i < BM * (PF_BK / 4); i += THREADS) {
            const int row = i / (PF_BK / 4), c4 = i % (PF_BK / 4);
            const bool ok = q0 + row < nq;
            cp_async16(a + row * PF_PITCH + c4 * 4, yhat + (size_t)(ok ? q0 + row : 0) * D + d0 + c4 * 4, ok);
        }
        for (int i = tid; i < BN * (PF_BK / 4); i += THREADS) {
            const int row = i / (PF_BK / 4), c4 = i % (PF_BK / 4);
            const bool ok = c0 + row < K;
            cp_async16(b + row * PF_PITCH + c4 * 4, chat + (size_t)(ok ? c0 + row : 0) * D + d0 + c4 * 4, ok);
        }
        asm volatile("cp.async.commit_group;");
    };

    const int nk = D / PF_BK;
    load_stage(0, 0);
    for (int kb = 0; kb < nk; kb++) {
        if (kb + 1 < nk) {
            load_stage((kb + 1) & 1, (kb + 1) * PF_BK);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        const float* a = sA + (kb & 1) * BM * PF_PITCH + (wm * TM * 16) * PF_PITCH;
        const float* b = sB + (kb & 1) * BN * PF_PITCH + (wn * TN * 8) * PF_PITCH;
#pragma unroll
        for (int k8 = 0; k8 < PF_BK; k8 += 8) {
            uint32_t af[TM][4], bf[TN][2];
#pragma unroll
            for (int i = 0; i < TM; i++) {
                const float* p = a + (i * 16 + g) * PF_PITCH + k8 + t;
                af[i][0] = __float_as_uint(p[0]);
                af[i][1] = __float_as_uint(p[8 * PF_PITCH]);
                af[i][2] = __float_as_uint(p[4]);
                af[i][3] = __float_as_uint(p[8 * PF_PITCH + 4]);
            }
#pragma unroll
            for (int j = 0; j < TN; j++) {
                const float* p = b + (j * 8 + g) * PF_PITCH + k8 + t;
                bf[j][0] = __float_as_uint(p[0]);
                bf[j][1] = __float_as_uint(p[4]);
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++)
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(acc[i][j][0]), "+f"(acc[i][j][1]), "+f"(acc[i][j][2]), "+f"(acc[i][j][3])
                                 : "r"(af[i][0]), "r"(af[i][1]), "r"(af[i][2]), "r"(af[i][3]), "r"(bf[j][0]), "r"(bf[j][1]));
        }
        __syncthreads();
    }
    // epilogue: A = ||c'||^2 - 2 S, or A -= 2 S for the cross terms of the 3xTF32 split   (c fragment: rows g / g+8, columns 2t / 2t+1)
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int c = c0 + wn * TN * 8 + j * 8 + 2 * t;
            const int qa = q0 + wm * TM * 16 + i * 16 + g, qb = qa + 8;
            if (c + 1 < K && (K & 1) == 0) {  // two adjacent columns per lane: 64-bit accesses
                const float2 n01 = accumulate ? make_float2(0.f, 0.f) : __ldg(reinterpret_cast<const float2*>(cnorm2 + c));
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int qq = h ? qb : qa;
                    if (qq < nq) {
                        float2* dst = reinterpret_cast<float2*>(A + (size_t)qq * K + c);
                        const float2 prev = accumulate ? *dst : n01;
                        *dst = make_float2(prev.x - 2.0f * acc[i][j][2 * h], prev.y - 2.0f * acc[i][j][2 * h + 1]);
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int cc = c + (e & 1), qq = (e & 2) ? qb : qa;
                    if (cc < K && qq < nq) {
                        float* dst = A + (size_t)qq * K + cc;
                        const float prev = accumulate ? *dst : __ldg(&cnorm2[cc]);
                        *dst = prev - 2.0f * acc[i][j][e];
                    }
                }
            }
        }
}

// ---- per query: threshold, candidates, exact distances of the candidates, the P nearest by (distance, id) --------------------
constexpr int PS_THREADS = 256;
constexpr int PS_CAP = 1024;  // candidate capacity

// m = alpha yn cn + 2^-21 (yn + cn)^2 + gamma max(a + yn^2, 0), every operation rounded up (a = the approximate key)
RQ_DEV float margin(float yn, float cn, float a, float alpha, float gamma) {
    const float s = __fadd_ru(yn, cn);
    const float e = fmaxf(__fadd_ru(a, __fmul_ru(yn, yn)), 0.0f);
    return __fadd_ru(__fadd_ru(__fmul_ru(__fmul_ru(yn, cn), alpha), __fmul_ru(__fmul_ru(s, s), 4.76837158203125e-07f)), __fmul_ru(e, gamma));
}

template <int PST>  // threads per query: 256, or 128 when the batch has several waves of CTAs (twice the queries per SM: the kernel is a chain of latencies)
__global__ void __launch_bounds__(PST) prefilter_select_kernel(
    const float* __restrict__ A, const float* __restrict__ ynorm, const float* __restrict__ cnorm, const float* __restrict__ cnorm_max,
    const float* __restrict__ y, const float* __restrict__ cent, int K, int P, int D, const uint32_t* __restrict__ offsets,
    const uint32_t* __restrict__ offsets_g, uint32_t* __restrict__ probe_ids, float* __restrict__ probe_dist,
    uint32_t* __restrict__ slot_local, uint32_t* __restrict__ q_words, uint32_t* __restrict__ q_pairs, uint32_t* __restrict__ q_p0,
    uint32_t* __restrict__ fallback_flag, int cap /* candidate capacity (even, <= PS_CAP; sizes the shared-memory arrays); tests lower it to exercise the fallback */, float alpha, float gamma) {
    extern __shared__ __align__(16) unsigned char ps_smem_raw[];
    float* sy = reinterpret_cast<float*>(ps_smem_raw);                            // D
    uint32_t* cid = reinterpret_cast<uint32_t*>(sy + D);                          // cap candidate ids
    unsigned long long* ckey = reinterpret_cast<unsigned long long*>(cid + cap);  // cap (okey(exact) << 32 | id)   (cap is even: 8-byte aligned)
    uint32_t* words = reinterpret_cast<uint32_t*>(ckey + cap);                    // P
    __shared__ uint32_t hist[256];
    __shared__ float s_red[2 * (PST / 32)];
    __shared__ uint32_t s_ncand, s_nle, s_p0, s_pairs;
    __shared__ uint32_t warp_tot[PST / 32 + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t q = blockIdx.x;
    const float* row = A + q * (size_t)K;
    const float yn = ynorm[q], cn_max = cnorm_max[0];

    for (int d = tid; d < D; d += PST) sy[d] = y[q * (size_t)D + d];
    // sample: 256 strided keys.  base = their minimum; hi = the largest of the minima of groups of G samples, i.e. roughly
    // the (ln(256/G) + 0.6) / G quantile of the keys -- a little above the P/K quantile the threshold has to reach, so that
    // only a small fraction of the keys enters the histogram (no hot bins) and the 255 bins below `hi` are narrow.
    int G = 64;
    while (G > 2 && (size_t)G * 4 * P > (size_t)K) G >>= 1;  // G ~ K / (4 P), a power of two in [2, 64]
    constexpr int SPT = 256 / PST;  // samples per thread
    float gmin = __ldg(&row[(size_t)(tid * SPT) * K / 256]);
#pragma unroll
    for (int i = 1; i < SPT; i++) gmin = fminf(gmin, __ldg(&row[(size_t)(tid * SPT + i) * K / 256]));
    const int Gt = max(1, G / SPT);  // group size in threads
    for (int o = 1; o < min(Gt, 32); o <<= 1) gmin = fminf(gmin, __shfl_xor_sync(FULL, gmin, o));
    float wmin = gmin, wmaxmin = gmin;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wmin = fminf(wmin, __shfl_xor_sync(FULL, wmin, o));
        wmaxmin = fmaxf(wmaxmin, __shfl_xor_sync(FULL, wmaxmin, o));
    }
    if (lane == 0) { s_red[warp] = wmin; s_red[PST / 32 + warp] = wmaxmin; }
    for (int i = tid; i < 256; i += PST) hist[i] = 0;
    if (tid == 0) { s_p0 = 0xffffffffu; s_pairs = 0; s_ncand = 0; s_nle = 0; }
    __syncthreads();
    float base = s_red[0], hi = -3.402823466e+38f;
#pragma unroll
    for (int w = 1; w < PST / 32; w++) base = fminf(base, s_red[w]);
    if (Gt == 64) {  // groups of two warps
#pragma unroll
        for (int w = 0; w < PST / 32; w += 2) hi = fmaxf(hi, fminf(s_red[w], s_red[w + 1]));
    } else {
#pragma unroll
        for (int w = 0; w < PST / 32; w++) hi = fmaxf(hi, s_red[PST / 32 + w]);
    }
    const bool vec4 = (K & 3) == 0;
    __shared__ float s_tau;
    __shared__ uint32_t s_bin0;
    float tau = 0.0f, kmin_true = 3.402823466e+38f;
    int refines = 0;
    for (int attempt = 0; attempt < 3;) {
        // pass A: 255-bin histogram of the keys below hi (keys below `base` fall into bin 0); the true minimum rides along
        const float scale = 255.0f / fmaxf(hi - base, 1e-30f);
        float vmin = 3.402823466e+38f;
        auto hadd = [&](float v) {
            vmin = fminf(vmin, v);
            if (v < hi) atomicAdd(&hist[max(0, min(254, (int)((v - base) * scale)))], 1u);
        };
        if (vec4) {  // 128-bit loads, several in flight per thread: the pass is a stream over K keys
            const float4* row4 = reinterpret_cast<const float4*>(row);
#pragma unroll 4
            for (int i4 = tid; i4 < K / 4; i4 += PST) {
                const float4 v = __ldg(&row4[i4]);
                hadd(v.x); hadd(v.y); hadd(v.z); hadd(v.w);
            }
        } else {
            for (int i = tid; i < K; i += PST) hadd(__ldg(&row[i]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmin = fminf(vmin, __shfl_xor_sync(FULL, vmin, o));
        if (lane == 0) s_red[warp] = vmin;
        __syncthreads();
        kmin_true = s_red[0];
#pragma unroll
        for (int w = 1; w < PST / 32; w++) kmin_true = fminf(kmin_true, s_red[w]);
        if (warp == 0) {  // first bin whose cumulative count reaches P -> tau = its upper edge (hi itself if none does)
            uint32_t c[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { c[j] = hist[lane * 8 + j]; s += c[j]; }
            uint32_t inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += v;
            }
            uint32_t run = inc - s;
            int found = -1;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                run += c[j];
                if (found < 0 && run >= (uint32_t)P) found = lane * 8 + j;
            }
            const uint32_t has = __ballot_sync(FULL, found >= 0);
            const int bsel = has ? __shfl_sync(FULL, found, __ffs(has) - 1) : 255;
            if (lane == 0) {
                s_tau = bsel >= 254 ? hi : fminf(hi, base + (float)(bsel + 1) / scale * 1.0001f);
                s_bin0 = bsel == 0 ? hist[0] : 0u;
            }
        }
        __syncthreads();
        tau = s_tau;
        if (s_bin0 > 2u * (uint32_t)P + 32u && refines < 2 && kmin_true < tau) {
            // the P-th key lies below the sampled range (P/K is smaller than the sample's resolution): the first bin holds far more
            // than P keys.  Zoom in: histogram [true minimum, upper edge of that bin] instead.
            refines++;
            hi = tau;
            base = kmin_true;
            __syncthreads();
            for (int i = tid; i < 256; i += PST) hist[i] = 0;
            __syncthreads();
            continue;
        }
        // pass B: candidates {A_c - m_c <= tau + m_max}, and the exact count of {A_c <= tau} that certifies tau.
        // The P keys <= tau have m_c <= mmax(tau); U bounds the reference's P-th distance (minus ||y'||^2); a candidate's own margin
        // is taken at its own key (the gamma term grows with the key), bounded by 1.5 mmax(U) for the quick reject.
        const float mmax = margin(yn, cn_max, tau, alpha, gamma);
        const float U = __fadd_ru(tau, mmax);
        const float Uq = __fadd_ru(U, __fmul_ru(margin(yn, cn_max, U, alpha, gamma), 1.5f));
        uint32_t nle = 0;
        auto consider = [&](float v, int i) {  // candidates are rare: a plain shared-memory append is cheap enough
            nle += v <= tau;
            if (v <= Uq && __fsub_rd(v, margin(yn, __ldg(&cnorm[i]), v, alpha, gamma)) <= U) {  // (quick reject first)
                const uint32_t pos = atomicAdd(&s_ncand, 1u);
                if (pos < (uint32_t)cap) cid[pos] = (uint32_t)i;
            }
        };
        if (vec4) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
#pragma unroll 4
            for (int i4 = tid; i4 < K / 4; i4 += PST) {
                const float4 v = __ldg(&row4[i4]);
                consider(v.x, 4 * i4); consider(v.y, 4 * i4 + 1); consider(v.z, 4 * i4 + 2); consider(v.w, 4 * i4 + 3);
            }
        } else {
            for (int i = tid; i < K; i += PST) consider(__ldg(&row[i]), i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nle += __shfl_xor_sync(FULL, nle, o);
        if (lane == 0 && nle) atomicAdd(&s_nle, nle);
        __syncthreads();
        if (s_nle >= (uint32_t)P) break;
        // fewer than P keys below hi (the sample was unlucky) or a bin edge a hair too low: widen and repeat (rare)
        __syncthreads();
        for (int i = tid; i < 256; i += PST) hist[i] = 0;
        if (tid == 0) { s_ncand = 0; s_nle = 0; }
        hi = attempt == 0 ? base + (hi - base) * 4.0f + 1e-30f : 3.402823466e+38f;
        attempt++;
        __syncthreads();
    }
    const uint32_t nc = s_ncand;
    if (nc > (uint32_t)cap || s_nle < (uint32_t)P) {  // cannot certify this query here: the batch is redone on the classic path
        if (tid == 0) atomicOr(fallback_flag, nc > (uint32_t)cap ? 1u : 2u);  // bit 0: too many candidates, bit 1: no valid threshold
        return;
    }
    if (tid == 0) atomicAdd(fallback_flag + 1, nc);  // statistics: candidates rechecked exactly (RABITQ_TRACE prints the mean)
    // exact distances of the candidates, order of simd::l2_squared_distance: four threads per candidate, two AVX lanes each
    // (packed f32x2, 64-bit loads of the centroid row; kernels.cuh l2_quad_global)
    {
        const int v4 = lane & 3, sub = tid >> 2;  // PST / 4 candidates per sweep
        for (uint32_t c0 = 0; c0 < nc; c0 += PST / 4) {
            const uint32_t ci = c0 + sub;
            const bool act = ci < nc;
            const uint32_t id = cid[act ? ci : 0];
            const float acc = l2_quad_global(cent + (size_t)id * D, sy, D, v4);  // diff = c - y (src/simd.rs:34-37)
            if (act && v4 == 0) ckey[ci] = ((unsigned long long)okey(acc) << 32) | id;
        }
    }
    __syncthreads();
    // the P smallest (key, id): rank by counting (candidates are few)
    for (uint32_t ci = tid; ci < nc; ci += PST) {
        const unsigned long long me = ckey[ci];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < nc; j++) rank += ckey[j] < me;
        if (rank < (uint32_t)P) {
            const uint32_t id = (uint32_t)(me & 0xffffffffu);
            probe_ids[q * P + rank] = id;
            probe_dist[q * P + rank] = okey_to_float((uint32_t)(me >> 32));
            const uint32_t n_c = offsets[id + 1] - offsets[id];
            words[rank] = (n_c + 31u) >> 5;
            const uint32_t n_g = offsets_g ? offsets_g[id + 1] - offsets_g[id] : n_c;
            if (n_g) { atomicAdd(&s_pairs, n_g); atomicMin(&s_p0, rank); }
        }
    }
    __syncthreads();
    const uint32_t total_words = block_exclusive_scan<PST>(words, P, warp_tot);
    for (int p = tid; p < P; p += PST) slot_local[q * P + p] = words[p];
    if (tid == 0) {
        q_pairs[q] = s_pairs;
        q_words[q] = total_words;
        q_p0[q] = s_p0 == 0xffffffffu ? 0u : s_p0;
    }
}

}  // namespace rq
