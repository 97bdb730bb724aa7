// rabitq_oracle.cpp -- CPU ORACLE for the IVF-RaBitQ query hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// The product path (rabitq_b200/csrc) never links, calls or falls back to anything here.
//
// What it is: a restatement, in C++ with the same AVX/AVX2/FMA intrinsics issued in the same
// order, of the reference kemingy/rabitq v0.2.2 query path (`RaBitQ::load_from_dir` +
// `RaBitQ::query`) plus the index builder (`from_path`) that produces its inputs.  Every
// function cites the reference file:line it follows (paths relative to /root/reference).
//
// PARITY UNPINNED: the reference is Rust, there is no Rust toolchain in this image, and the
// reference ships no tests, golden vectors or fixtures for this path.
// The oracle is therefore pinned only by (i) the reference's own AVX2-vs-scalar twin
// implementations restated below and cross-checked in tests/, (ii) an independent float64
// numpy model of the estimator (tests/model_f64.py), (iii) recall against brute force, (iv) a second,
// plain-Python restatement of both rerankers replayed over the visit-order trace, (v) numpy emulations of
// the fp32 evaluation ORDER of project / the centroid distances / the estimator that must match bit for bit
// (tests/test_oracle_cpu.py; DESIGN.md section 2 lists all seven pins).
//
// Build: see oracle/Makefile  (g++ -O2 -mavx2 -mfma -ffp-contract=off; never -ffast-math,
// never -mpopcnt: the reference's scalar popcount path is compiled without the popcnt target
// feature, src/simd.rs:324, so `count_ones` lowers to the SWAR sequence restated here).

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---- src/consts.rs:4-12 --------------------------------------------------------------------
constexpr float DEFAULT_X_DOT_PRODUCT = 0.8f;
constexpr float EPSILON = 1.9f;
constexpr uint32_t THETA_LOG_DIM = 4;
constexpr float SCALAR = 1.0f / (float(1u << THETA_LOG_DIM) - 1.0f);
constexpr size_t WINDOW_SIZE = 12;

// ---- src/metrics.rs:7-16,44-61,65 ----------------------------------------------------------
struct Metrics {
    std::atomic<uint64_t> rough{0}, precise{0}, query{0}, miss{0};
};
Metrics METRICS;

// ---- src/ord32.rs:12-26 --------------------------------------------------------------------
inline int32_t ord32_from_f32(float x) {
    int32_t bits;
    std::memcpy(&bits, &x, 4);
    uint32_t mask = uint32_t(bits >> 31) >> 1;
    return bits ^ int32_t(mask);
}
inline float ord32_to_f32(int32_t v) {
    uint32_t mask = uint32_t(v >> 31) >> 1;
    int32_t bits = v ^ int32_t(mask);
    float x;
    std::memcpy(&x, &bits, 4);
    return x;
}

// `u64::count_ones` without the popcnt target feature (src/simd.rs:324,336,342): LLVM's
// generic ctpop lowering.
inline uint32_t count_ones_swar(uint64_t x) {
    x = x - ((x >> 1) & 0x5555555555555555ull);
    x = (x & 0x3333333333333333ull) + ((x >> 2) & 0x3333333333333333ull);
    x = (x + (x >> 4)) & 0x0f0f0f0f0f0f0f0full;
    return uint32_t((x * 0x0101010101010101ull) >> 56);
}

// ---- src/simd.rs:52-63 / 292-303: horizontal reduce ((s0+s4)+(s1+s5))+((s2+s6)+(s3+s7)) ------
inline float reduce_f32_256(__m256 acc) {
    __m256 combined = _mm256_add_ps(acc, _mm256_permute2f128_ps(acc, acc, 1));
    combined = _mm256_hadd_ps(combined, combined);
    combined = _mm256_hadd_ps(combined, combined);
    return _mm256_cvtss_f32(combined);
}

// ---- src/simd.rs:14-73 ---------------------------------------------------------------------
float simd_l2_squared_distance(const float* lhs, const float* rhs, size_t len) {
    __m256 sum = _mm256_setzero_ps();
    for (size_t i = 0; i < len / 16; i++) {
        __m256 vx = _mm256_loadu_ps(lhs), vy = _mm256_loadu_ps(rhs);
        lhs += 8; rhs += 8;
        __m256 diff = _mm256_sub_ps(vx, vy);
        sum = _mm256_fmadd_ps(diff, diff, sum);
        vx = _mm256_loadu_ps(lhs); vy = _mm256_loadu_ps(rhs);
        lhs += 8; rhs += 8;
        diff = _mm256_sub_ps(vx, vy);
        sum = _mm256_fmadd_ps(diff, diff, sum);
    }
    for (size_t i = 0; i < (len & 15) / 8; i++) {
        __m256 vx = _mm256_loadu_ps(lhs), vy = _mm256_loadu_ps(rhs);
        lhs += 8; rhs += 8;
        __m256 diff = _mm256_sub_ps(vx, vy);
        sum = _mm256_fmadd_ps(diff, diff, sum);
    }
    float res = reduce_f32_256(sum);
    for (size_t i = 0; i < (len & 7); i++) {
        float residual = *lhs - *rhs;
        res += residual * residual;
        lhs++; rhs++;
    }
    return res;
}

// ---- src/simd.rs:257-314 -------------------------------------------------------------------
float simd_vector_dot_product(const float* lhs, const float* rhs, size_t len) {
    __m256 acc = _mm256_setzero_ps();
    for (size_t i = 0; i < len / 16; i++) {
        acc = _mm256_fmadd_ps(_mm256_loadu_ps(lhs), _mm256_loadu_ps(rhs), acc);
        lhs += 8; rhs += 8;
        acc = _mm256_fmadd_ps(_mm256_loadu_ps(lhs), _mm256_loadu_ps(rhs), acc);
        lhs += 8; rhs += 8;
    }
    for (size_t i = 0; i < (len & 15) / 8; i++) {
        acc = _mm256_fmadd_ps(_mm256_loadu_ps(lhs), _mm256_loadu_ps(rhs), acc);
        lhs += 8; rhs += 8;
    }
    float sum = reduce_f32_256(acc);
    for (size_t i = 0; i < (len & 7); i++) {
        sum += *lhs * *rhs;
        lhs++; rhs++;
    }
    return sum;
}

// ---- src/simd.rs:117-173 -------------------------------------------------------------------
void simd_min_max_residual(float* res, const float* x, const float* y, size_t len, float* out_min,
                           float* out_max) {
    __m256 min8 = _mm256_set1_ps(std::numeric_limits<float>::max());
    __m256 max8 = _mm256_set1_ps(std::numeric_limits<float>::lowest());
    float mn = std::numeric_limits<float>::max(), mx = std::numeric_limits<float>::lowest();
    float buf[8];
    size_t rest = len & 7;
    for (size_t i = 0; i < len / 8; i++) {
        __m256 r = _mm256_sub_ps(_mm256_loadu_ps(x), _mm256_loadu_ps(y));
        _mm256_storeu_ps(res, r);
        x += 8; y += 8; res += 8;
        min8 = _mm256_min_ps(min8, r);
        max8 = _mm256_max_ps(max8, r);
    }
    _mm256_storeu_ps(buf, min8);
    for (float v : buf) if (v < mn) mn = v;
    _mm256_storeu_ps(buf, max8);
    for (float v : buf) if (v > mx) mx = v;
    for (size_t i = 0; i < rest; i++) {
        *res = *x - *y;
        if (*res < mn) mn = *res;
        if (*res > mx) mx = *res;
        res++; x++; y++;
    }
    *out_min = mn;
    *out_max = mx;
}

// ---- src/simd.rs:185-247 (round-to-nearest-even via cvtps_epi32; the bias is NOT used) --------
uint32_t simd_scalar_quantize(uint8_t* quantized, const float* vec, size_t len, float lower_bound,
                              float multiplier) {
    __m256 lower = _mm256_set1_ps(lower_bound);
    __m256 scalar = _mm256_set1_ps(multiplier);
    __m256i sum256 = _mm256_setzero_si256();
    const __m256i mask = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    size_t rest = len & 7;
    uint8_t* qp = quantized;
    for (size_t i = 0; i < len / 8; i++) {
        __m256 v = _mm256_loadu_ps(vec);
        __m256i q = _mm256_cvtps_epi32(_mm256_mul_ps(_mm256_sub_ps(v, lower), scalar));
        sum256 = _mm256_add_epi32(sum256, q);
        __m256i shuffled = _mm256_shuffle_epi8(q, mask);
        uint64_t packed = uint64_t(uint32_t(_mm256_extract_epi32(shuffled, 0))) |
                          (uint64_t(uint32_t(_mm256_extract_epi32(shuffled, 4))) << 32);
        std::memcpy(qp, &packed, 8);
        qp += 8;
        vec += 8;
    }
    __m256i combined = _mm256_add_epi32(sum256, _mm256_permute2f128_si256(sum256, sum256, 1));
    combined = _mm256_hadd_epi32(combined, combined);
    combined = _mm256_hadd_epi32(combined, combined);
    uint32_t sum = uint32_t(_mm256_cvtsi256_si32(combined));
    for (size_t i = 0; i < rest; i++) {
        uint8_t q = uint8_t(std::round((*vec - lower_bound) * multiplier));
        quantized[len - rest + i] = q;
        sum += q;
        vec++;
    }
    return sum;
}

// ---- src/simd.rs:83-107 --------------------------------------------------------------------
void simd_vector_binarize_query(const uint8_t* vec, size_t length, uint64_t* binary) {
    for (size_t i = 0; i < length; i += 32) {
        __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(vec + i));
        v = _mm256_slli_epi32(v, 4);
        for (size_t j = 0; j < THETA_LOG_DIM; j++) {
            uint64_t m = uint64_t(uint32_t(_mm256_movemask_epi8(v)));
            size_t shift = i & 32;
            binary[(3 - j) * (length >> 6) + (i >> 6)] |= m << shift;
            v = _mm256_slli_epi32(v, 1);
        }
    }
}

// ---- src/simd.rs:326-384 -------------------------------------------------------------------
inline __m256i mm256_popcnt_epi64(__m256i x) {
    const __m256i lut = _mm256_setr_epi8(0, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4,
                                         0, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4);
    const __m256i mask = _mm256_set1_epi8(15);
    const __m256i zero = _mm256_setzero_si256();
    __m256i low = _mm256_and_si256(x, mask);
    __m256i high = _mm256_and_si256(_mm256_srli_epi64(x, 4), mask);
    low = _mm256_shuffle_epi8(lut, low);
    high = _mm256_shuffle_epi8(lut, high);
    return _mm256_sad_epu8(_mm256_add_epi8(low, high), zero);
}

uint32_t simd_binary_dot_product(const uint64_t* lhs, const uint64_t* rhs, size_t n) {
    uint32_t sum = 0;
    size_t length = n / 4;
    if (length == 0) {
        for (size_t i = 0; i < n; i++) sum += count_ones_swar(lhs[i] & rhs[i]);
        return sum;
    }
    size_t rest = n & 3;
    for (size_t i = 0; i < rest; i++) sum += count_ones_swar(lhs[4 * length + i] & rhs[4 * length + i]);
    __m256i sum256 = _mm256_setzero_si256();
    const __m256i* xp = reinterpret_cast<const __m256i*>(lhs);
    const __m256i* yp = reinterpret_cast<const __m256i*>(rhs);
    for (size_t i = 0; i < length; i++) {
        __m256i a = _mm256_and_si256(_mm256_loadu_si256(xp + i), _mm256_loadu_si256(yp + i));
        sum256 = _mm256_add_epi64(sum256, mm256_popcnt_epi64(a));
    }
    __m128i xa = _mm_add_epi64(_mm256_castsi256_si128(sum256), _mm256_extracti128_si256(sum256, 1));
    sum += uint32_t(_mm_cvtsi128_si32(_mm_add_epi64(xa, _mm_shuffle_epi32(xa, 78))));
    return sum;
}

// ---- src/utils.rs:113-135 (AVX2 branch) ----------------------------------------------------
inline uint32_t asymmetric_binary_dot_product(const uint64_t* x, const uint64_t* y, size_t w) {
    uint32_t res = 0;
    for (uint32_t i = 0; i < THETA_LOG_DIM; i++) {
        res += simd_binary_dot_product(x, y, w) << i;
        y += w;
    }
    return res;
}

// ===== scalar twins (never the parity target; kept for cross-checks) ======
// src/utils.rs:90-97
void raw_vector_binarize_query(const uint8_t* vec, size_t length, uint64_t* binary) {
    for (size_t j = 0; j < THETA_LOG_DIM; j++)
        for (size_t i = 0; i < length; i++)
            binary[(i + j * length) / 64] |= uint64_t((vec[i] >> j) & 1) << (i % 64);
}
// src/utils.rs:101-107
uint32_t raw_binary_dot_product(const uint64_t* x, const uint64_t* y, size_t n) {
    uint32_t res = 0;
    for (size_t i = 0; i < n; i++) res += count_ones_swar(x[i] & y[i]);
    return res;
}
// src/utils.rs:155-168
void raw_min_max(float* res, const float* x, const float* y, size_t len, float* mn_out, float* mx_out) {
    float mn = std::numeric_limits<float>::max(), mx = std::numeric_limits<float>::lowest();
    for (size_t i = 0; i < len; i++) {
        res[i] = x[i] - y[i];
        if (res[i] < mn) mn = res[i];
        if (res[i] > mx) mx = res[i];
    }
    *mn_out = mn;
    *mx_out = mx;
}
// src/utils.rs:194-209 (truncate + bias: deliberately different rounding from the AVX2 twin)
uint32_t raw_scalar_quantize(uint8_t* quantized, const float* vec, const float* bias, size_t len,
                             float lower_bound, float multiplier) {
    uint32_t sum = 0;
    for (size_t i = 0; i < len; i++) {
        float f = (vec[i] - lower_bound) * multiplier + bias[i];
        // Rust `as u8` saturates and maps NaN to 0.
        uint8_t q = (f != f) ? 0 : (f <= 0.f ? 0 : (f >= 255.f ? 255 : uint8_t(f)));
        quantized[i] = q;
        sum += q;
    }
    return sum;
}

// ===== vecs IO (src/utils.rs:280-364): record = u32 count then count elements, little endian ====
template <typename T>
bool read_vecs_file(const std::string& path, std::vector<std::vector<T>>& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    uint32_t dim;
    while (std::fread(&dim, 4, 1, f) == 1) {
        std::vector<T> v(dim);
        if (dim && std::fread(v.data(), sizeof(T), dim, f) != dim) { std::fclose(f); return false; }
        out.emplace_back(std::move(v));
    }
    std::fclose(f);
    return true;
}
template <typename T>
bool write_vecs_file(const std::string& path, const T* data, size_t rows, size_t cols) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    uint32_t c = uint32_t(cols);
    for (size_t r = 0; r < rows; r++) {
        std::fwrite(&c, 4, 1, f);
        std::fwrite(data + r * cols, sizeof(T), cols, f);
    }
    std::fclose(f);
    return true;
}

// ===== index (src/rabitq.rs:21-32,57-68) ========================================================
struct Factor { float factor_ip, factor_ppc, error_bound, center_distance_square; };

struct Index {
    uint32_t dim = 0;      // padded D, multiple of 64
    size_t n = 0, k = 0;
    std::vector<float> base;        // n x D, vector u contiguous  (faer D x N col-major, rabitq.rs:110-112)
    std::vector<float> orth_cols;   // D x D, column i of P contiguous: orth_cols[i*D + r] = P[r][i]
    std::vector<float> orth_rows;   // D x D as stored on disk:        orth_rows[r*D + i] = P[r][i]
    std::vector<float> centroids;   // k x D, rotated, centroid c contiguous (faer D x K col-major)
    std::vector<float> rand_bias;   // D, only used by the non-AVX2 twin (rabitq.rs:119)
    bool use_raw_quantize = false;  // true: scalar_quantize takes its non-AVX2 branch (utils.rs:222-228) with rand_bias
    std::vector<uint32_t> offsets;  // k+1
    std::vector<uint32_t> map_ids;  // n
    std::vector<uint64_t> x_binary_vec;  // n * D/64
    std::vector<Factor> factors;    // n
};

thread_local std::string g_err;

// ---- src/rabitq.rs:84-125 ------------------------------------------------------------------
Index* load_from_dir(const std::string& dir) {
    auto idx = new Index();
    std::vector<std::vector<float>> orth, cent, fac, base;
    std::vector<std::vector<uint32_t>> offs;
    std::vector<std::vector<uint64_t>> xb;
    if (!read_vecs_file(dir + "/orthogonal.fvecs", orth) || orth.empty()) { g_err = "open orthogonal error"; delete idx; return nullptr; }
    if (!read_vecs_file(dir + "/centroids.fvecs", cent) || cent.empty()) { g_err = "open centroids error"; delete idx; return nullptr; }
    if (!read_vecs_file(dir + "/offsets_ids.ivecs", offs) || offs.empty()) { g_err = "open offsets_ids error"; delete idx; return nullptr; }
    if (!read_vecs_file(dir + "/factors.fvecs", fac)) { g_err = "open factors error"; delete idx; return nullptr; }
    if (!read_vecs_file(dir + "/x_binary_vec.u64vecs", xb)) { g_err = "open x_binary_vec error"; delete idx; return nullptr; }
    if (!read_vecs_file(dir + "/base.fvecs", base)) { g_err = "read vecs error"; delete idx; return nullptr; }
    size_t D = orth.size();
    if (D % 64 != 0) { g_err = "assertion failed: dim % 64 == 0"; delete idx; return nullptr; }
    idx->dim = uint32_t(D);
    idx->orth_rows.resize(D * D);
    idx->orth_cols.resize(D * D);
    for (size_t r = 0; r < D; r++)
        for (size_t i = 0; i < D; i++) {
            idx->orth_rows[r * D + i] = orth[r][i];
            idx->orth_cols[i * D + r] = orth[r][i];
        }
    // centroids.fvecs holds D records of K floats (matrix D x K written row-wise, rabitq.rs:133).
    size_t K = cent[0].size();
    idx->k = K;
    idx->centroids.resize(K * D);
    for (size_t d = 0; d < D; d++)
        for (size_t c = 0; c < K; c++) idx->centroids[c * D + d] = cent[d][c];
    idx->offsets = offs.front();
    idx->map_ids = offs.back();
    std::vector<float> flat;
    for (auto& v : fac) flat.insert(flat.end(), v.begin(), v.end());
    idx->factors.resize(flat.size() / 4);
    std::memcpy(idx->factors.data(), flat.data(), idx->factors.size() * sizeof(Factor));
    for (auto& v : xb) idx->x_binary_vec.insert(idx->x_binary_vec.end(), v.begin(), v.end());
    idx->n = base.size();
    idx->base.resize(idx->n * D);
    for (size_t u = 0; u < idx->n; u++) std::memcpy(&idx->base[u * D], base[u].data(), D * 4);
    idx->rand_bias.assign(D, 0.5f);
    return idx;
}

// ---- src/rabitq.rs:128-156 -----------------------------------------------------------------
bool dump_to_dir(const Index* idx, const std::string& dir) {
    size_t D = idx->dim, K = idx->k;
    if (!write_vecs_file(dir + "/base.fvecs", idx->base.data(), idx->n, D)) return false;
    if (!write_vecs_file(dir + "/orthogonal.fvecs", idx->orth_rows.data(), D, D)) return false;
    std::vector<float> cdk(D * K);
    for (size_t d = 0; d < D; d++)
        for (size_t c = 0; c < K; c++) cdk[d * K + c] = idx->centroids[c * D + d];
    if (!write_vecs_file(dir + "/centroids.fvecs", cdk.data(), D, K)) return false;
    {
        FILE* f = std::fopen((dir + "/offsets_ids.ivecs").c_str(), "wb");
        if (!f) return false;
        uint32_t c = uint32_t(idx->offsets.size());
        std::fwrite(&c, 4, 1, f);
        std::fwrite(idx->offsets.data(), 4, c, f);
        c = uint32_t(idx->map_ids.size());
        std::fwrite(&c, 4, 1, f);
        std::fwrite(idx->map_ids.data(), 4, c, f);
        std::fclose(f);
    }
    if (!write_vecs_file(dir + "/factors.fvecs", reinterpret_cast<const float*>(idx->factors.data()), 1, idx->factors.size() * 4)) return false;
    if (!write_vecs_file(dir + "/x_binary_vec.u64vecs", idx->x_binary_vec.data(), 1, idx->x_binary_vec.size())) return false;
    return true;
}

// Deterministic orthogonal matrix: Q of a seeded standard-normal D x D matrix (src/utils.rs:16-20 draws
// from an unseeded thread_rng and uses faer's QR; any orthogonal P is a valid index).
struct SplitMix { uint64_t s; uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
                  double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); } };
void gen_orthogonal(size_t D, uint64_t seed, std::vector<float>& rows) {
    SplitMix rng{seed};
    std::vector<double> a(D * D);  // column-major columns a[j*D + r]
    for (size_t i = 0; i < D * D; i += 2) {
        double u1 = rng.uniform(), u2 = rng.uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        double r = std::sqrt(-2.0 * std::log(u1));
        a[i] = r * std::cos(6.283185307179586 * u2);
        if (i + 1 < D * D) a[i + 1] = r * std::sin(6.283185307179586 * u2);
    }
    // modified Gram-Schmidt, two passes
    for (size_t j = 0; j < D; j++) {
        double* cj = &a[j * D];
        for (int pass = 0; pass < 2; pass++)
            for (size_t p = 0; p < j; p++) {
                const double* cp = &a[p * D];
                double dot = 0;
                for (size_t r = 0; r < D; r++) dot += cj[r] * cp[r];
                for (size_t r = 0; r < D; r++) cj[r] -= dot * cp[r];
            }
        double nrm = 0;
        for (size_t r = 0; r < D; r++) nrm += cj[r] * cj[r];
        nrm = std::sqrt(nrm);
        for (size_t r = 0; r < D; r++) cj[r] /= nrm;
    }
    rows.resize(D * D);
    for (size_t r = 0; r < D; r++)
        for (size_t i = 0; i < D; i++) rows[r * D + i] = float(a[i * D + r]);
}

// ---- src/rabitq.rs:159-265 (index build; produces the inputs of the hot path) -----------------
// base: n x len row-major (original space), centroids: k x len.  P (D x D, row r = P[r,:]) may be null.
Index* build_from_arrays(const float* base_in, size_t n, size_t len, const float* cent_in, size_t k,
                         const float* P, uint64_t seed, int nthreads) {
    size_t D = (len + 63) / 64 * 64;  // rabitq.rs:167-179
    auto idx = new Index();
    idx->dim = uint32_t(D);
    idx->n = n;
    idx->k = k;
    std::vector<float> base(n * D, 0.f), cent(k * D, 0.f);
    for (size_t i = 0; i < n; i++) std::memcpy(&base[i * D], base_in + i * len, len * 4);
    for (size_t i = 0; i < k; i++) std::memcpy(&cent[i * D], cent_in + i * len, len * 4);
    if (P) idx->orth_rows.assign(P, P + D * D);
    else gen_orthogonal(D, seed, idx->orth_rows);  // rabitq.rs:182
    idx->orth_cols.resize(D * D);
    for (size_t r = 0; r < D; r++)
        for (size_t i = 0; i < D; i++) idx->orth_cols[i * D + r] = idx->orth_rows[r * D + i];
    idx->rand_bias.assign(D, 0.5f);

    // rabitq.rs:188-189: x_projected = base * P, centroids = centroids * P (faer matmul in the reference;
    // the summation order of the builder needs no parity, it only defines the shared index).
    std::vector<float> xp(n * D);
    idx->centroids.resize(k * D);
    auto project_rows = [&](const float* src, float* dst, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++)
            for (size_t c = 0; c < D; c++)
                dst[i * D + c] = simd_vector_dot_product(src + i * D, &idx->orth_cols[c * D], D);
    };
    project_rows(cent.data(), idx->centroids.data(), 0, k);
    if (nthreads < 1) nthreads = 1;
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++)
            th.emplace_back([&, t] { project_rows(base.data(), xp.data(), n * t / nthreads, n * (t + 1) / nthreads); });
        for (auto& x : th) x.join();
    }

    float dim_sqrt = std::sqrt(float(D));  // rabitq.rs:192
    std::vector<uint32_t> label(n);
    std::vector<float> min_dist(n), x_c_distance(n), x_dot_product(n), sum_sign(n);
    std::vector<Factor> factors(n);
    size_t W = D / 64;
    std::vector<uint64_t> codes(n * W, 0);
    auto prep = [&](size_t lo, size_t hi) {
        std::vector<float> r(D);
        for (size_t i = lo; i < hi; i++) {
            const float* x = &xp[i * D];
            // utils.rs:261-277 kmeans_nearest_cluster: strict <, first minimum wins
            float md = std::numeric_limits<float>::max();
            size_t ml = 0;
            for (size_t j = 0; j < k; j++) {
                float d = simd_l2_squared_distance(&idx->centroids[j * D], x, D);
                if (d < md) { md = d; ml = j; }
            }
            label[i] = uint32_t(ml);
            min_dist[i] = md;
            const float* c = &idx->centroids[ml * D];
            float ss = 0.f, dot = 0.f, ssign = 0.f;
            for (size_t d = 0; d < D; d++) {
                r[d] = x[d] - c[d];               // rabitq.rs:205
                ss += r[d] * r[d];
                bool pos = r[d] > 0.0f;           // utils.rs:53-67
                if (pos) codes[i * W + d / 64] |= 1ull << (d % 64);
                float s = pos ? 1.0f : -1.0f;
                dot += r[d] * s;
                ssign += s;
            }
            x_c_distance[i] = std::sqrt(ss);      // rabitq.rs:206 norm_l2
            factors[i].center_distance_square = x_c_distance[i] * x_c_distance[i];  // :207 powi(2)
            float norm = x_c_distance[i] * dim_sqrt;                                 // :210
            x_dot_product[i] = std::isnormal(norm) ? dot / norm : DEFAULT_X_DOT_PRODUCT;  // :211-215
            sum_sign[i] = ssign;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back([&, t] { prep(n * t / nthreads, n * (t + 1) / nthreads); });
        for (auto& x : th) x.join();
    }
    // rabitq.rs:218-229
    float error_base = 2.0f * EPSILON / std::sqrt(float(D) - 1.0f);
    for (size_t i = 0; i < n; i++) {
        float x_c_over_ip = x_c_distance[i] / x_dot_product[i];
        Factor& f = factors[i];
        f.error_bound = error_base * std::sqrt(x_c_over_ip * x_c_over_ip - f.center_distance_square);
        f.factor_ip = -2.0f / dim_sqrt * x_c_over_ip;
        f.factor_ppc = f.factor_ip * sum_sign[i];
    }
    // rabitq.rs:231-252: stable sort by distance to centroid inside each cluster, prefix offsets, permute
    std::vector<std::vector<uint32_t>> lists(k);
    for (size_t i = 0; i < n; i++) lists[label[i]].push_back(uint32_t(i));
    idx->offsets.assign(k + 1, 0);
    idx->map_ids.clear();
    idx->map_ids.reserve(n);
    for (size_t c = 0; c < k; c++) {
        auto& v = lists[c];
        std::stable_sort(v.begin(), v.end(), [&](uint32_t a, uint32_t b) { return min_dist[a] < min_dist[b]; });
        idx->offsets[c + 1] = idx->offsets[c] + uint32_t(v.size());
        idx->map_ids.insert(idx->map_ids.end(), v.begin(), v.end());
    }
    idx->base.resize(n * D);
    idx->x_binary_vec.resize(n * W);
    idx->factors.resize(n);
    for (size_t i = 0; i < n; i++) {
        uint32_t s = idx->map_ids[i];
        std::memcpy(&idx->base[i * D], &base[size_t(s) * D], D * 4);  // base stays UNROTATED, :245-247
        std::memcpy(&idx->x_binary_vec[i * W], &codes[size_t(s) * W], W * 8);
        idx->factors[i] = factors[s];
    }
    return idx;
}

// ===== rerankers (src/rerank.rs) ================================================================
struct HeapItem { int32_t key; uint32_t id; };

// std::collections::BinaryHeap<(Ord32, AlwaysEqual<u32>)>: max-heap on key, ids never compared
// (src/ord32.rs:43-66).  push/pop restate the std algorithm (sift_up; pop = swap last into the root,
// sift_down_to_bottom, sift_up) so that equal-key eviction follows the same path.
struct RustBinaryHeap {
    std::vector<HeapItem> data;
    void sift_up(size_t start, size_t pos) {
        HeapItem e = data[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (e.key <= data[parent].key) break;
            data[pos] = data[parent];
            pos = parent;
        }
        data[pos] = e;
    }
    void push(HeapItem it) {
        size_t old = data.size();
        data.push_back(it);
        sift_up(0, old);
    }
    void pop() {
        HeapItem item = data.back();
        data.pop_back();
        if (data.empty()) return;
        std::swap(item, data[0]);
        size_t end = data.size(), pos = 0;
        HeapItem e = data[0];
        size_t child = 1;
        while (child <= (end >= 2 ? end - 2 : 0) && end >= 2) {
            if (data[child].key <= data[child + 1].key) child++;
            data[pos] = data[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            data[pos] = data[child];
            pos = child;
        }
        data[pos] = e;
        sift_up(0, pos);
    }
};

struct PairTrace {  // optional per-candidate record of what the reranker did
    float* exact = nullptr;    // exact distance if computed else NaN
    uint8_t* action = nullptr; // 0 = filtered, 1 = exact computed, 2 = pushed
};

// src/rerank.rs:61-114
struct HeapReRanker {
    float threshold = std::numeric_limits<float>::max();
    size_t topk;
    RustBinaryHeap heap;
    const float* query;
    HeapReRanker(const float* q, size_t k) : topk(k), query(q) { heap.data.reserve(k + 1); }
    void rank_batch(const std::pair<float, uint32_t>* rough, size_t cnt, const Index& ix, uint64_t* precise_out,
                    PairTrace* tr, size_t tr_off) {
        uint64_t precise = 0;
        size_t D = ix.dim;
        for (size_t t = 0; t < cnt; t++) {
            float r = rough[t].first;
            uint32_t u = rough[t].second;
            if (tr && tr->action) { tr->action[tr_off + t] = 0; tr->exact[tr_off + t] = std::numeric_limits<float>::quiet_NaN(); }
            if (r < threshold) {
                float accurate = simd_l2_squared_distance(&ix.base[size_t(u) * D], query, D);
                precise++;
                if (tr && tr->action) { tr->action[tr_off + t] = 1; tr->exact[tr_off + t] = accurate; }
                if (accurate < threshold) {
                    if (tr && tr->action) tr->action[tr_off + t] = 2;
                    heap.push({ord32_from_f32(accurate), ix.map_ids[u]});
                    if (heap.data.size() > topk) heap.pop();
                    if (heap.data.size() == topk) threshold = ord32_to_f32(heap.data[0].key);
                }
            }
        }
        *precise_out += precise;
    }
    size_t get_result(float* dist, uint32_t* ids) const {
        for (size_t i = 0; i < heap.data.size(); i++) { dist[i] = ord32_to_f32(heap.data[i].key); ids[i] = heap.data[i].id; }
        return heap.data.size();
    }
};

// src/rerank.rs:117-176
struct HeuristicReRanker {
    float threshold = std::numeric_limits<float>::max();
    float recent_max_accurate = std::numeric_limits<float>::lowest();
    size_t topk, count = 0;
    std::vector<std::pair<float, uint32_t>> array;
    const float* query;
    HeuristicReRanker(const float* q, size_t k) : topk(k), query(q) {}
    void rank_batch(const std::pair<float, uint32_t>* rough, size_t cnt, const Index& ix, uint64_t* precise_out,
                    PairTrace* tr, size_t tr_off) {
        uint64_t precise = 0;
        size_t D = ix.dim;
        for (size_t t = 0; t < cnt; t++) {
            float r = rough[t].first;
            uint32_t u = rough[t].second;
            if (tr && tr->action) { tr->action[tr_off + t] = 0; tr->exact[tr_off + t] = std::numeric_limits<float>::quiet_NaN(); }
            if (r < threshold) {
                float accurate = simd_l2_squared_distance(&ix.base[size_t(u) * D], query, D);
                precise++;
                if (tr && tr->action) { tr->action[tr_off + t] = 1; tr->exact[tr_off + t] = accurate; }
                if (accurate < threshold) {
                    if (tr && tr->action) tr->action[tr_off + t] = 2;
                    array.emplace_back(accurate, ix.map_ids[u]);
                    count++;
                    recent_max_accurate = std::max(recent_max_accurate, accurate);
                    if (count >= WINDOW_SIZE) {
                        threshold = recent_max_accurate;
                        count = 0;
                        recent_max_accurate = std::numeric_limits<float>::lowest();
                    }
                }
            }
        }
        *precise_out += precise;
    }
    size_t get_result(float* dist, uint32_t* ids) const {
        size_t length = std::min(topk, array.size());
        if (length == 0) return 0;  // the reference would panic on `length - 1` underflow
        auto res = array;
        // select_nth_unstable_by + truncate: the set is defined, the order is not; return ascending.
        std::stable_sort(res.begin(), res.end(), [](auto& a, auto& b) { return ord32_from_f32(a.first) < ord32_from_f32(b.first); });
        for (size_t i = 0; i < length; i++) { dist[i] = res[i].first; ids[i] = res[i].second; }
        return length;
    }
};

struct Trace {  // all caller-allocated; any pointer may be null
    float* y = nullptr;            // D
    float* centroid_dist = nullptr;  // K
    uint32_t* probe_ids = nullptr;   // P
    float* probe_dist = nullptr;     // P
    float* lo = nullptr;             // P
    float* delta = nullptr;          // P
    uint32_t* sum = nullptr;         // P
    uint64_t* planes = nullptr;      // P * 4W
    uint8_t* quantized = nullptr;    // P * D
    float* rough = nullptr;          // pairs (concatenated in visit order)
    uint32_t* abdp = nullptr;        // pairs
    uint32_t* pair_pos = nullptr;    // pairs: sorted position j
    PairTrace pt;                    // pairs
    size_t pair_capacity = 0;
    size_t pairs = 0;                // out
    uint64_t precise = 0, roughc = 0;  // out
};

// ---- src/rabitq.rs:268-333 + 336-367 ---------------------------------------------------------
size_t query_impl(const Index& ix, const float* query, size_t len, size_t probe, size_t topk, bool heuristic,
                  float* out_dist, uint32_t* out_ids, Trace* tr, uint64_t* rough_cnt, uint64_t* precise_cnt) {
    size_t D = ix.dim, W = D / 64, K = ix.k;
    std::vector<float> query_vec(D, 0.f);                       // rabitq.rs:277-280
    std::memcpy(query_vec.data(), query, len * 4);
    std::vector<float> y(D);
    for (size_t i = 0; i < D; i++)                              // utils.rs:237-258 project
        y[i] = simd_vector_dot_product(query_vec.data(), &ix.orth_cols[i * D], D);
    if (tr && tr->y) std::memcpy(tr->y, y.data(), D * 4);
    std::vector<std::pair<float, size_t>> lists(K);
    for (size_t i = 0; i < K; i++)                              // rabitq.rs:285-293
        lists[i] = {simd_l2_squared_distance(&ix.centroids[i * D], y.data(), D), i};
    if (tr && tr->centroid_dist) for (size_t i = 0; i < K; i++) tr->centroid_dist[i] = lists[i].first;
    size_t length = std::min(probe, K);
    // rabitq.rs:294-297: select_nth_unstable_by + truncate + stable sort by total_cmp.  Which of several
    // EQUAL distances survives the unstable select is unspecified in the reference; the oracle (and the
    // CUDA path) break such ties by the smaller centroid index.
    std::sort(lists.begin(), lists.end(), [](auto& a, auto& b) {
        int32_t ka = ord32_from_f32(a.first), kb = ord32_from_f32(b.first);
        return ka != kb ? ka < kb : a.second < b.second;
    });
    lists.resize(length);

    HeapReRanker heap_rr(query_vec.data(), topk);               // rabitq.rs:299: UNROTATED padded query
    HeuristicReRanker heur_rr(query_vec.data(), topk);
    std::vector<float> residual(D);
    std::vector<uint8_t> quantized(D);
    std::vector<std::pair<float, uint32_t>> rough_distances;
    std::vector<uint64_t> binary_vec(W * THETA_LOG_DIM);
    uint64_t precise = 0, roughc = 0;
    size_t pair_off = 0;
    for (size_t p = 0; p < length; p++) {
        float dist = lists[p].first;
        size_t i = lists[p].second;
        float lower_bound, upper_bound;
        simd_min_max_residual(residual.data(), y.data(), &ix.centroids[i * D], D, &lower_bound, &upper_bound);
        float delta = (upper_bound - lower_bound) * SCALAR;     // rabitq.rs:307
        float one_over_delta = 1.0f / delta;                    // :308 recip()
        uint32_t scalar_sum = ix.use_raw_quantize   // utils.rs:213-232: AVX2 host -> simd, otherwise the raw twin with the bias
                                  ? raw_scalar_quantize(quantized.data(), residual.data(), ix.rand_bias.data(), D, lower_bound, one_over_delta)
                                  : simd_scalar_quantize(quantized.data(), residual.data(), D, lower_bound, one_over_delta);
        std::fill(binary_vec.begin(), binary_vec.end(), 0);
        simd_vector_binarize_query(quantized.data(), D, binary_vec.data());
        if (tr) {
            if (tr->probe_ids) tr->probe_ids[p] = uint32_t(i);
            if (tr->probe_dist) tr->probe_dist[p] = dist;
            if (tr->lo) tr->lo[p] = lower_bound;
            if (tr->delta) tr->delta[p] = delta;
            if (tr->sum) tr->sum[p] = scalar_sum;
            if (tr->planes) std::memcpy(tr->planes + p * 4 * W, binary_vec.data(), 4 * W * 8);
            if (tr->quantized) std::memcpy(tr->quantized + p * D, quantized.data(), D);
        }
        // rabitq.rs:336-367 calculate_rough_distance
        float ssum = float(scalar_sum);                         // `scalar_sum as f32`, :322
        float dist_sqrt = std::sqrt(dist);
        for (uint32_t j = ix.offsets[i]; j < ix.offsets[i + 1]; j++) {
            const Factor& f = ix.factors[j];
            uint32_t ab = asymmetric_binary_dot_product(&ix.x_binary_vec[size_t(j) * W], binary_vec.data(), W);
            float rough = f.center_distance_square + dist + lower_bound * f.factor_ppc +
                          (2.0f * float(ab) - ssum) * f.factor_ip * delta - f.error_bound * dist_sqrt;
            rough_distances.emplace_back(rough, j);
            if (tr) {
                size_t t = pair_off + rough_distances.size() - 1;
                if (t < tr->pair_capacity) {
                    if (tr->rough) tr->rough[t] = rough;
                    if (tr->abdp) tr->abdp[t] = ab;
                    if (tr->pair_pos) tr->pair_pos[t] = j;
                }
            }
        }
        PairTrace* ptr = (tr && tr->pt.action && pair_off + rough_distances.size() <= tr->pair_capacity) ? &tr->pt : nullptr;
        if (heuristic) heur_rr.rank_batch(rough_distances.data(), rough_distances.size(), ix, &precise, ptr, pair_off);
        else heap_rr.rank_batch(rough_distances.data(), rough_distances.size(), ix, &precise, ptr, pair_off);
        roughc += rough_distances.size();                       // rerank.rs:105
        pair_off += rough_distances.size();
        rough_distances.clear();
    }
    METRICS.precise.fetch_add(precise, std::memory_order_relaxed);
    METRICS.rough.fetch_add(roughc, std::memory_order_relaxed);
    METRICS.query.fetch_add(1, std::memory_order_relaxed);      // rabitq.rs:331
    if (tr) { tr->pairs = pair_off; tr->precise = precise; tr->roughc = roughc; }
    if (rough_cnt) *rough_cnt += roughc;
    if (precise_cnt) *precise_cnt += precise;
    return heuristic ? heur_rr.get_result(out_dist, out_ids) : heap_rr.get_result(out_dist, out_ids);
}

}  // namespace

// ===== C interface for ctypes (tests / bench cpu_baseline only) =================================
extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_load_from_dir(const char* dir) { return load_from_dir(dir); }
int orc_dump_to_dir(const void* idx, const char* dir) { return dump_to_dir(static_cast<const Index*>(idx), dir) ? 0 : 1; }
void* orc_build(const float* base, size_t n, size_t len, const float* centroids, size_t k, const float* P,
                uint64_t seed, int nthreads) {
    return build_from_arrays(base, n, len, centroids, k, P, seed, nthreads);
}
// Adopt already-built arrays (same logical content as the six files).  centroids: K x D, rotated, each
// centroid contiguous.  orthogonal: D x D, row r = P[r,:].
void* orc_from_arrays(uint32_t dim, size_t n, size_t k, const float* base, const float* orthogonal,
                      const float* centroids, const uint32_t* offsets, const uint32_t* map_ids,
                      const uint64_t* codes, const float* factors) {
    if (dim % 64) { g_err = "assertion failed: dim % 64 == 0"; return nullptr; }
    auto idx = new Index();
    size_t D = dim;
    idx->dim = dim; idx->n = n; idx->k = k;
    idx->base.assign(base, base + n * D);
    idx->orth_rows.assign(orthogonal, orthogonal + D * D);
    idx->orth_cols.resize(D * D);
    for (size_t r = 0; r < D; r++) for (size_t i = 0; i < D; i++) idx->orth_cols[i * D + r] = orthogonal[r * D + i];
    idx->centroids.assign(centroids, centroids + k * D);
    idx->offsets.assign(offsets, offsets + k + 1);
    idx->map_ids.assign(map_ids, map_ids + n);
    idx->x_binary_vec.assign(codes, codes + n * (D / 64));
    idx->factors.resize(n);
    std::memcpy(idx->factors.data(), factors, n * sizeof(Factor));
    idx->rand_bias.assign(D, 0.5f);
    return idx;
}
void orc_free(void* idx) { delete static_cast<Index*>(idx); }
// bias != NULL: query() quantises like the reference on a host WITHOUT AVX2 (scalar_quantize_raw with this rand_bias);
// NULL: back to the AVX2 branch.
void orc_set_raw_bias(void* idx, const float* bias) {
    Index* ix = static_cast<Index*>(idx);
    ix->use_raw_quantize = bias != nullptr;
    if (bias) ix->rand_bias.assign(bias, bias + ix->dim);
}

uint32_t orc_dim(const void* idx) { return static_cast<const Index*>(idx)->dim; }
size_t orc_n(const void* idx) { return static_cast<const Index*>(idx)->n; }
size_t orc_k(const void* idx) { return static_cast<const Index*>(idx)->k; }
const float* orc_base(const void* idx) { return static_cast<const Index*>(idx)->base.data(); }
const float* orc_orthogonal(const void* idx) { return static_cast<const Index*>(idx)->orth_rows.data(); }
const float* orc_centroids(const void* idx) { return static_cast<const Index*>(idx)->centroids.data(); }
const uint32_t* orc_offsets(const void* idx) { return static_cast<const Index*>(idx)->offsets.data(); }
const uint32_t* orc_map_ids(const void* idx) { return static_cast<const Index*>(idx)->map_ids.data(); }
const uint64_t* orc_codes(const void* idx) { return static_cast<const Index*>(idx)->x_binary_vec.data(); }
const float* orc_factors(const void* idx) { return reinterpret_cast<const float*>(static_cast<const Index*>(idx)->factors.data()); }

// RaBitQ::query (src/rabitq.rs:268-274).  Result in the reference's order (heap-internal for the heap
// reranker).  Returns the number of results, or -1 when the reference would panic on the dim assert.
int orc_query(const void* idx, const float* q, size_t len, size_t probe, size_t topk, int heuristic,
              float* out_dist, uint32_t* out_ids) {
    const Index& ix = *static_cast<const Index*>(idx);
    if (ix.dim != (len + 63) / 64 * 64) { g_err = "assertion `left == right` failed (dim)"; return -1; }
    return int(query_impl(ix, q, len, probe, topk, heuristic != 0, out_dist, out_ids, nullptr, nullptr, nullptr));
}

// The CLI loop (crates/cli/src/main.rs:69-75) over nq queries on `nthreads` threads (1 = the reference).
// out_* are nq x topk, out_count nq.  counters[0..1] += rough, precise.  Returns seconds spent in query().
double orc_query_batch(const void* idx, const float* queries, size_t nq, size_t len, size_t probe, size_t topk,
                       int heuristic, int nthreads, float* out_dist, uint32_t* out_ids, uint32_t* out_count,
                       uint64_t* counters) {
    const Index& ix = *static_cast<const Index*>(idx);
    if (ix.dim != (len + 63) / 64 * 64) { g_err = "assertion `left == right` failed (dim)"; return -1.0; }
    if (nthreads < 1) nthreads = 1;
    std::vector<uint64_t> rc(nthreads, 0), pc(nthreads, 0);
    std::vector<double> secs(nthreads, 0.0);
    auto work = [&](int t) {
        size_t lo = nq * t / nthreads, hi = nq * (t + 1) / nthreads;
        std::vector<float> d(topk + 1);
        std::vector<uint32_t> ids(topk + 1);
        for (size_t i = lo; i < hi; i++) {
            timespec a, b;
            clock_gettime(CLOCK_MONOTONIC, &a);
            size_t c = query_impl(ix, queries + i * len, len, probe, topk, heuristic != 0, d.data(), ids.data(), nullptr, &rc[t], &pc[t]);
            clock_gettime(CLOCK_MONOTONIC, &b);
            secs[t] += double(b.tv_sec - a.tv_sec) + 1e-9 * double(b.tv_nsec - a.tv_nsec);
            if (out_count) out_count[i] = uint32_t(c);
            for (size_t j = 0; j < c; j++) {
                if (out_dist) out_dist[i * topk + j] = d[j];
                if (out_ids) out_ids[i * topk + j] = ids[j];
            }
        }
    };
    timespec w0, w1;
    clock_gettime(CLOCK_MONOTONIC, &w0);
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    clock_gettime(CLOCK_MONOTONIC, &w1);
    if (counters) for (int t = 0; t < nthreads; t++) { counters[0] += rc[t]; counters[1] += pc[t]; }
    if (nthreads == 1) return secs[0];  // sum of per-query times, exactly the CLI's `total_time`
    return double(w1.tv_sec - w0.tv_sec) + 1e-9 * double(w1.tv_nsec - w0.tv_nsec);
}

// One query with every intermediate exposed.  Arrays are caller-allocated; pass NULL to skip one.
// Returns the result count; *pairs_out = number of (query, candidate) pairs visited.
int orc_query_trace(const void* idx, const float* q, size_t len, size_t probe, size_t topk, int heuristic,
                    float* out_dist, uint32_t* out_ids, float* y, float* centroid_dist, uint32_t* probe_ids,
                    float* probe_dist, float* lo, float* delta, uint32_t* sum, uint64_t* planes, uint8_t* quantized,
                    size_t pair_capacity, float* rough, uint32_t* abdp, uint32_t* pair_pos, float* exact,
                    uint8_t* action, uint64_t* pairs_out, uint64_t* precise_out) {
    const Index& ix = *static_cast<const Index*>(idx);
    if (ix.dim != (len + 63) / 64 * 64) { g_err = "assertion `left == right` failed (dim)"; return -1; }
    Trace tr;
    tr.y = y; tr.centroid_dist = centroid_dist; tr.probe_ids = probe_ids; tr.probe_dist = probe_dist;
    tr.lo = lo; tr.delta = delta; tr.sum = sum; tr.planes = planes; tr.quantized = quantized;
    tr.rough = rough; tr.abdp = abdp; tr.pair_pos = pair_pos; tr.pt.exact = exact; tr.pt.action = action;
    if (!exact || !action) { tr.pt.exact = nullptr; tr.pt.action = nullptr; }
    tr.pair_capacity = pair_capacity;
    int c = int(query_impl(ix, q, len, probe, topk, heuristic != 0, out_dist, out_ids, &tr, nullptr, nullptr));
    if (pairs_out) *pairs_out = tr.pairs;
    if (precise_out) *precise_out = tr.precise;
    return c;
}

// src/metrics.rs:30-41 order: query, rough, precise, miss
void orc_metrics(uint64_t out[4]) {
    out[0] = METRICS.query.load(); out[1] = METRICS.rough.load(); out[2] = METRICS.precise.load(); out[3] = METRICS.miss.load();
}
void orc_metrics_reset() { METRICS.query = 0; METRICS.rough = 0; METRICS.precise = 0; METRICS.miss = 0; }

// ---- unit kernels, AVX2 path and scalar twins -------------------------------------------------
float orc_l2_squared_distance(const float* a, const float* b, size_t n) { return simd_l2_squared_distance(a, b, n); }
float orc_vector_dot_product(const float* a, const float* b, size_t n) { return simd_vector_dot_product(a, b, n); }
void orc_min_max_residual(float* res, const float* x, const float* y, size_t n, float* mn, float* mx) { simd_min_max_residual(res, x, y, n, mn, mx); }
void orc_min_max_raw(float* res, const float* x, const float* y, size_t n, float* mn, float* mx) { raw_min_max(res, x, y, n, mn, mx); }
uint32_t orc_scalar_quantize(uint8_t* q, const float* v, size_t n, float lo, float mul) { return simd_scalar_quantize(q, v, n, lo, mul); }
uint32_t orc_scalar_quantize_raw(uint8_t* q, const float* v, const float* bias, size_t n, float lo, float mul) { return raw_scalar_quantize(q, v, bias, n, lo, mul); }
void orc_vector_binarize_query(const uint8_t* v, size_t n, uint64_t* out) { simd_vector_binarize_query(v, n, out); }
void orc_vector_binarize_query_raw(const uint8_t* v, size_t n, uint64_t* out) { raw_vector_binarize_query(v, n, out); }
uint32_t orc_binary_dot_product(const uint64_t* a, const uint64_t* b, size_t n) { return simd_binary_dot_product(a, b, n); }
uint32_t orc_binary_dot_product_raw(const uint64_t* a, const uint64_t* b, size_t n) { return raw_binary_dot_product(a, b, n); }
uint32_t orc_asymmetric_binary_dot_product(const uint64_t* x, const uint64_t* y, size_t w) { return asymmetric_binary_dot_product(x, y, w); }
int32_t orc_ord32_from_f32(float x) { return ord32_from_f32(x); }
float orc_ord32_to_f32(int32_t v) { return ord32_to_f32(v); }
float orc_scalar_const() { return SCALAR; }
void orc_gen_orthogonal(size_t D, uint64_t seed, float* out_rows) {
    std::vector<float> r;
    gen_orthogonal(D, seed, r);
    std::memcpy(out_rows, r.data(), D * D * 4);
}
// Heap exposed for the replay unit tests: feed (exact) values with the rerank.rs:92-101 logic.
int orc_heap_replay(const float* accurate, const uint32_t* ids, size_t n, size_t topk, float* out_dist, uint32_t* out_ids) {
    RustBinaryHeap h;
    float thr = std::numeric_limits<float>::max();
    for (size_t i = 0; i < n; i++) {
        if (accurate[i] < thr) {
            h.push({ord32_from_f32(accurate[i]), ids[i]});
            if (h.data.size() > topk) h.pop();
            if (h.data.size() == topk) thr = ord32_to_f32(h.data[0].key);
        }
    }
    for (size_t i = 0; i < h.data.size(); i++) { out_dist[i] = ord32_to_f32(h.data[i].key); out_ids[i] = h.data[i].id; }
    return int(h.data.size());
}

}  // extern "C"
