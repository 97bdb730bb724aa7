// build_index.cuh -- index training on the device: RaBitQ::from_path (reference src/rabitq.rs:159-265) and
// dump_to_dir (src/rabitq.rs:128-156).  This is the step immediately BEFORE the query path (SURVEY.md section 8f
// rank 1); it produces the arrays of `struct RaBitQ` directly in HBM.  The reference draws P from an unseeded RNG
// and multiplies with faer, so there is no bit-parity target for the builder; the formulas are the reference's:
//   label  = first nearest rotated centroid by simd::l2_squared_distance          (utils.rs:261-277)
//   r      = x*P - c_label ; bits = r > 0 ; s = +-1                               (rabitq.rs:205-209, utils.rs:53-67)
//   |r|, cds = |r|^2, x_dot = <r,s> / (|r| sqrt(D)) if that is normal else 0.8    (rabitq.rs:206-215)
//   t = |r| / x_dot ; err = 2*1.9/sqrt(D-1) * sqrt(t^2 - cds) ; ip = -2/sqrt(D) * t ; ppc = ip * sum(s)  (rabitq.rs:220-228)
//   clusters in id order, inside a cluster ascending distance to the centroid, stable                   (rabitq.rs:231-252)
#pragma once

#include <cub/device/device_radix_sort.cuh>

#include "kernels.cuh"

namespace rq {

// ---- P = Q factor of a seeded standard-normal matrix (utils.rs:16-20), classical Gram-Schmidt applied twice, fp64 ----
RQ_DEV unsigned long long splitmix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void gauss_fill_kernel(double* __restrict__ A, size_t n, unsigned long long seed) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long a = splitmix64(seed ^ (2 * i + 1)), b = splitmix64(seed ^ (2 * i + 2) ^ 0x5851f42d4c957f2dull);
    double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (b >> 11) * (1.0 / 9007199254740992.0);
    A[i] = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

// dots[p] = <A[:,p], A[:,j]> for p < j   (A row-major D x D; thread = p, coalesced over p)
__global__ void gs_dots_kernel(const double* __restrict__ A, int D, int j, double* __restrict__ dots) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= j) return;
    double s = 0.0;
    for (int r = 0; r < D; r++) s += A[(size_t)r * D + p] * A[(size_t)r * D + j];
    dots[p] = s;
}
// A[r][j] -= sum_p dots[p] * A[r][p]
__global__ void gs_update_kernel(double* __restrict__ A, int D, int j, const double* __restrict__ dots) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= D) return;
    double s = 0.0;
    const double* row = A + (size_t)r * D;
    for (int p = 0; p < j; p++) s += dots[p] * row[p];
    A[(size_t)r * D + j] -= s;
}
__global__ void __launch_bounds__(256) gs_normalize_kernel(double* __restrict__ A, int D, int j) {
    __shared__ double red[256];
    double s = 0.0;
    for (int r = threadIdx.x; r < D; r += 256) { double v = A[(size_t)r * D + j]; s += v * v; }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const double inv = 1.0 / sqrt(red[0]);
    for (int r = threadIdx.x; r < D; r += 256) A[(size_t)r * D + j] *= inv;
}
__global__ void f64_to_f32_kernel(const double* __restrict__ a, float* __restrict__ b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) b[i] = (float)a[i];
}

// ---- assignment: first minimum of a row of squared distances (strict <, utils.rs:271) ---------------------------------
__global__ void argmin_rows_kernel(const float* __restrict__ dist, size_t rows, int K, uint32_t* __restrict__ label,
                                   float* __restrict__ min_dist) {
    const int lane = threadIdx.x & 31;
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* d = dist + row * (size_t)K;
    float best = 3.402823466e+38f;  // f32::MAX: a distance must be strictly smaller to win
    int bi = 0x7fffffff;
    for (int c = lane; c < K; c += 32) {
        const float v = d[c];
        if (v < best) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(FULL, best, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
        label[row] = bi == 0x7fffffff ? 0u : (uint32_t)bi;  // nothing below f32::MAX -> label 0 like the reference
        min_dist[row] = best;
    }
}

// ---- per-vector code + Factor, one warp per vector --------------------------------------------------------------------
__global__ void __launch_bounds__(128) encode_kernel(const float* __restrict__ xp /* rows x D, rotated */,
                                                     const float* __restrict__ cent /* K x D, rotated */,
                                                     const uint32_t* __restrict__ label, const float* __restrict__ min_dist,
                                                     size_t rows, int D, uint32_t* __restrict__ codes /* rows x D/32 */,
                                                     float4* __restrict__ factors, unsigned long long* __restrict__ sort_key) {
    const int lane = threadIdx.x & 31;
    const size_t row = blockIdx.x * (size_t)4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const uint32_t lab = label[row];
    const float* x = xp + row * (size_t)D;
    const float* c = cent + (size_t)lab * D;
    float ss = 0.0f, dot = 0.0f, ssign = 0.0f;
    for (int g = 0; g < D / 32; g++) {
        const float r = __fsub_rn(x[g * 32 + lane], c[g * 32 + lane]);
        const bool pos = r > 0.0f;
        const uint32_t w = __ballot_sync(FULL, pos);
        if (lane == 0) codes[row * (size_t)(D / 32) + g] = w;  // bit (i % 64) of u64 word i / 64 == bit (i % 32) of u32 word i / 32
        const float s = pos ? 1.0f : -1.0f;
        ss = __fadd_rn(ss, __fmul_rn(r, r));
        dot = __fadd_rn(dot, __fmul_rn(r, s));
        ssign += s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ss = __fadd_rn(ss, __shfl_xor_sync(FULL, ss, o));
        dot = __fadd_rn(dot, __shfl_xor_sync(FULL, dot, o));
        ssign += __shfl_xor_sync(FULL, ssign, o);
    }
    if (lane == 0) {
        const float dim_sqrt = __fsqrt_rn((float)D);
        const float error_base = __fdiv_rn(2.0f * 1.9f, __fsqrt_rn((float)D - 1.0f));
        const float norm = __fsqrt_rn(ss);
        const float cds = __fmul_rn(norm, norm);
        const float nrm = __fmul_rn(norm, dim_sqrt);
        const bool normal = isfinite(nrm) && fabsf(nrm) >= 1.17549435e-38f;
        const float x_dot = normal ? __fdiv_rn(dot, nrm) : 0.8f;
        const float t = __fdiv_rn(norm, x_dot);
        const float err = __fmul_rn(error_base, __fsqrt_rn(__fsub_rn(__fmul_rn(t, t), cds)));
        const float ip = __fmul_rn(__fdiv_rn(-2.0f, dim_sqrt), t);
        factors[row] = make_float4(ip, __fmul_rn(ip, ssign), err, cds);
        // (label, distance) ascending; squared distances are >= 0 so their bit patterns sort like the values
        sort_key[row] = ((unsigned long long)lab << 32) | (unsigned long long)__float_as_uint(min_dist[row]);
    }
}

__global__ void iota_kernel(uint32_t* __restrict__ p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

__global__ void label_histogram_kernel(const uint32_t* __restrict__ label, size_t n, uint32_t* __restrict__ counts) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[label[i]], 1u);
}

__global__ void __launch_bounds__(1024) offsets_scan_kernel(const uint32_t* __restrict__ counts, int K, uint32_t* __restrict__ offsets) {
    __shared__ uint32_t wt[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (K + 1023) / 1024;
    const int lo = min(K, tid * per), hi = min(K, lo + per);
    uint32_t s = 0;
    for (int c = lo; c < hi; c++) s += counts[c];
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wt[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wt[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(FULL, wi, o);
            if (lane >= o) wi += v;
        }
        wt[lane] = wi - w;
        if (lane == 31) wt[32] = wi;
    }
    __syncthreads();
    uint32_t run = wt[warp] + inc - s;
    for (int c = lo; c < hi; c++) {
        offsets[c] = run;
        run += counts[c];
    }
    if (tid == 0) offsets[K] = wt[32];
}

// cluster-sorted copies: base (zero-padded to D, UNROTATED, rabitq.rs:245-247), codes, factors, map_ids
__global__ void __launch_bounds__(128) permute_kernel(const float* __restrict__ base_in, int len, int D, const uint32_t* __restrict__ perm,
                                                      const uint32_t* __restrict__ codes_in, const float4* __restrict__ factors_in,
                                                      size_t n, float* __restrict__ base_out, uint32_t* __restrict__ codes_out,
                                                      float4* __restrict__ factors_out, uint32_t* __restrict__ map_ids) {
    const int lane = threadIdx.x & 31;
    const size_t i = blockIdx.x * (size_t)4 + (threadIdx.x >> 5);
    if (i >= n) return;
    const size_t src = perm[i];
    for (int d = lane; d < D; d += 32) base_out[i * (size_t)D + d] = d < len ? base_in[src * (size_t)len + d] : 0.0f;
    for (int w = lane; w < D / 32; w += 32) codes_out[i * (size_t)(D / 32) + w] = codes_in[src * (size_t)(D / 32) + w];
    if (lane == 0) {
        factors_out[i] = factors_in[src];
        map_ids[i] = (uint32_t)src;
    }
}

__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {  // out[c][r] = in[r][c]
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    out[(size_t)c * rows + r] = in[i];
}

}  // namespace rq
