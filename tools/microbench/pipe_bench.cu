// pipe_bench.cu -- issue-rate microbenchmark for the instructions the code scan lives on (POPC, LOP3, IADD3, IMAD)
// on sm_100a.  Prints thread-instructions per clock per SM.  Harness only.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u; y[i] = seed ^ (i * 77u); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) {  // POPC only (dependent chain per i, ILP independent chains)
                x[i] = __popc(x[i]) + 0x55555555u * 0;  // popc feeds itself
                asm volatile("" : "+r"(x[i]));
            } else if (MODE == 1) {  // LOP3 only
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(seed));
            } else if (MODE == 2) {  // AND + POPC + ADD (the plain scan inner op)
                uint32_t a = x[i] & y[i];
                uint32_t p = __popc(a);
                y[i] += p;
                asm volatile("" : "+r"(y[i]));
                x[i] = x[i] * 1u + 0;  // keep
            } else if (MODE == 3) {  // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            } else if (MODE == 4) {  // IMAD
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(seed));
            } else if (MODE == 5) {  // 2x LOP3 + 1 POPC  (CSA-heavy mix)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(y[i]) : "r"(x[i]), "r"(seed));
                uint32_t p = __popc(x[i]);
                asm volatile("" : "+r"(p));
                y[i] ^= p;
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double inst_per_iter, int blocks_per_sm) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * blocks_per_sm;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaMalloc(&cyc, blocks * 8);
    k<MODE><<<blocks, 256>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += h[i];
    avg /= blocks;
    double thread_inst_per_sm = (double)blocks_per_sm * 256 * ITERS * ILP * inst_per_iter;
    printf("%-28s blocks/SM=%d  cycles=%.0f  -> %.1f thread-inst/clk/SM (counted inst/iter=%.0f)  wall %.3f ms => %.2f GHz eff\n", name,
           blocks_per_sm, avg, thread_inst_per_sm / avg, inst_per_iter, ms, avg / (ms * 1e6));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int b : {2, 4, 8}) {
        run<0>("POPC", 1, b);
        run<1>("LOP3", 1, b);
        run<2>("AND+POPC+IADD (3 inst)", 3, b);
        run<3>("IADD", 1, b);
        run<4>("IMAD", 1, b);
        run<5>("2xLOP3+POPC+XOR (4 inst)", 4, b);
    }
    return 0;
}
