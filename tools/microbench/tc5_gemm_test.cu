// Standalone check of approx_gemm_tc5_kernel against a plain fp32/fp64 reference and against the mma.sync kernel's timing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rabitq_b200/csrc -o /tmp/tc5_test tools/microbench/tc5_gemm_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "prefilter.cuh"
#include "tc5_gemm.cuh"
using namespace rq;

__global__ void ref_kernel(const float* y, const float* c, const float* cn2, int nq, int K, int D, float* A) {
    const int cc = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
    if (cc >= K) return;
    double s = 0;
    for (int d = 0; d < D; d++) s += (double)y[(size_t)q * D + d] * (double)c[(size_t)cc * D + d];
    A[(size_t)q * K + cc] = (float)((double)cn2[cc] - 2.0 * s);
}
__global__ void fill_kernel(float* p, size_t n, uint32_t seed) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = (uint32_t)i * 2654435761u + seed;
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    p[i] = to_tf32(((float)(x & 0xffffff) / 8388608.0f - 1.0f));
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

int run(int nq, int K, int D) {
    float *y, *c, *cn2, *A, *R;
    CK(cudaMalloc(&y, (size_t)nq * D * 4)); CK(cudaMalloc(&c, (size_t)K * D * 4)); CK(cudaMalloc(&cn2, K * 4));
    CK(cudaMalloc(&A, (size_t)nq * K * 4)); CK(cudaMalloc(&R, (size_t)nq * K * 4));
    fill_kernel<<<(unsigned)(((size_t)nq * D + 255) / 256), 256>>>(y, (size_t)nq * D, 1u);
    fill_kernel<<<(unsigned)(((size_t)K * D + 255) / 256), 256>>>(c, (size_t)K * D, 77u);
    fill_kernel<<<(K + 255) / 256, 256>>>(cn2, K, 5u);
    CK(cudaMemset(A, 0xff, (size_t)nq * K * 4));
    ref_kernel<<<dim3((K + 127) / 128, nq), 128>>>(y, c, cn2, nq, K, D, R);
    CUtensorMap tmy, tmc;
    if (tc5_make_tmap(&tmy, y, nq, D, TC5_BM) || tc5_make_tmap(&tmc, c, K, D, TC5_BN)) { printf("tensor map encode failed\n"); return 1; }
    CK(cudaFuncSetAttribute(approx_gemm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC5_SMEM_BYTES));
    dim3 grid((K + TC5_BN - 1) / TC5_BN, (nq + TC5_BM - 1) / TC5_BM);
    approx_gemm_tc5_kernel<<<grid, TC5_THREADS, TC5_SMEM_BYTES>>>(tmy, tmc, cn2, nq, K, D, A, 0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> ha((size_t)nq * K), hr((size_t)nq * K);
    CK(cudaMemcpy(ha.data(), A, ha.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), R, hr.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; size_t bad = 0;
    for (size_t i = 0; i < ha.size(); i++) {
        double e = fabs((double)ha[i] - (double)hr[i]);
        if (!(e <= 1e-3 * D / 128.0 + 1e-3)) { if (bad < 5) printf("  mismatch at q=%zu c=%zu: %f vs %f\n", i / K, i % K, ha[i], hr[i]); bad++; }
        if (e > maxerr) maxerr = e;
    }
    // accumulate mode: A += -2 <y, c>  ->  A = cn2 - 4 <y, c>
    approx_gemm_tc5_kernel<<<grid, TC5_THREADS, TC5_SMEM_BYTES>>>(tmy, tmc, cn2, nq, K, D, A, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ha.data(), A, ha.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<float> hcn(K);
    CK(cudaMemcpy(hcn.data(), cn2, K * 4, cudaMemcpyDeviceToHost));
    size_t bad2 = 0;
    for (size_t i = 0; i < ha.size(); i++) {
        double want = 2.0 * hr[i] - hcn[i % K];
        if (!(fabs(ha[i] - want) <= 2e-3 * D / 128.0 + 2e-3)) bad2++;
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms5 = 0, msold = 0;
    for (int it = 0; it < 3; it++) approx_gemm_tc5_kernel<<<grid, TC5_THREADS, TC5_SMEM_BYTES>>>(tmy, tmc, cn2, nq, K, D, A, 0);
    cudaEventRecord(e0);
    for (int it = 0; it < 20; it++) approx_gemm_tc5_kernel<<<grid, TC5_THREADS, TC5_SMEM_BYTES>>>(tmy, tmc, cn2, nq, K, D, A, 0);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms5, e0, e1);
    CK(cudaFuncSetAttribute(approx_gemm_tf32_kernel<128, 128, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * PF_PITCH * 4));
    dim3 g2((K + 127) / 128, (nq + 127) / 128);
    for (int it = 0; it < 3; it++) approx_gemm_tf32_kernel<128, 128, 2, 4><<<g2, 256, 2 * 256 * PF_PITCH * 4>>>(y, c, cn2, nq, K, D, A, 0);
    cudaEventRecord(e0);
    for (int it = 0; it < 20; it++) approx_gemm_tf32_kernel<128, 128, 2, 4><<<g2, 256, 2 * 256 * PF_PITCH * 4>>>(y, c, cn2, nq, K, D, A, 0);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&msold, e0, e1);
    printf("nq=%d K=%d D=%d: max |err| %.3g, mismatches %zu, accumulate-mode mismatches %zu; tcgen05 %.4f ms, mma.sync %.4f ms per launch\n", nq, K, D,
           maxerr, bad, bad2, ms5 / 20, msold / 20);
    cudaFree(y); cudaFree(c); cudaFree(cn2); cudaFree(A); cudaFree(R);
    return bad || bad2 ? 2 : 0;
}

int main() {
    int rc = 0;
    rc |= run(1000, 1024, 960);
    rc |= run(10000, 4096, 128);
    rc |= run(333, 1000, 64);
    rc |= run(10000, 16384, 128);
    printf(rc ? "FAILED\n" : "OK\n");
    return rc;
}
