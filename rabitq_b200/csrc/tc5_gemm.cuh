// tc5_gemm.cuh -- the prefilter's approximate-key GEMM on the 5th-generation tensor cores (src/rabitq.rs:283-293 is what the keys
// stand in for; prefilter.cuh explains how they are used and why the result stays bit-identical to the reference).
//
//   A[q][c] = ||c'||^2 - 2 <y^[q], c^[c]>      y^ : nq x D, c^ : K x D, both fp32 values already rounded to TF32, d-contiguous
//
// Blackwell form of approx_gemm_tf32_kernel: operands are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a ring of
// shared-memory stages, ONE elected thread issues tcgen05.mma kind::tf32 (M = 128 queries x N = 128 centroids x K = 8 per
// instruction) with the accumulator tile in TENSOR MEMORY, tcgen05.commit releases a stage / announces the finished tile through
// mbarriers, and four epilogue warps read the accumulators back with tcgen05.ld (32 lanes x 32 columns per instruction), finish
// the key and store it.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (warp w reads TMEM lanes 32 (w % 4) ..).
#pragma once

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through the runtime, the library links no libcuda)
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace rq {

constexpr int TC5_BM = 128, TC5_BN = 128, TC5_BK = 32, TC5_STAGES = 3;  // 97 KB of shared memory: two CTAs per SM, one's epilogue under the other's main loop
constexpr int TC5_STAGE_BYTES = (TC5_BM + TC5_BN) * TC5_BK * 4;        // 32 KB: 128 query rows + 128 centroid rows of 128 bytes
constexpr int TC5_SMEM_BYTES = TC5_STAGES * TC5_STAGE_BYTES + 1024 + 256;  // + alignment slack + barriers
constexpr int TC5_THREADS = 192;

typedef CUresult (*tc5_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tc5_encode_fn tc5_encoder() {
    static tc5_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tc5_encode_fn>(p);
    }
    return fn;
}

// tensor map of a row-major [rows][D] fp32 matrix, box = 32 floats (one 128-byte swizzle row) x box_rows; rows beyond the matrix read as 0
inline int tc5_make_tmap(CUtensorMap* tm, const float* ptr, size_t rows, size_t D, uint32_t box_rows) {
    tc5_encode_fn enc = tc5_encoder();
    if (!enc) return -1;
    const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)D * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TC5_BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

RQ_DEV void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
RQ_DEV void tma_load_2d(void* dst_smem, const CUtensorMap* tm, int c_inner, int c_row, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(c_inner), "r"(c_row), "r"(smem_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor of a K-major operand tile in the 128-byte-swizzle layout TMA writes: rows of 128 bytes, 8-row
// groups 1024 bytes apart (stride byte offset), leading byte offset unused by swizzled K-major layouts (1), descriptor version 1
RQ_DEV uint64_t tc5_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128, dense, no negation
constexpr uint32_t TC5_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC5_BN >> 3) << 17) | ((uint32_t)(TC5_BM >> 4) << 24);

RQ_DEV void tc5_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(TC5_IDESC), "r"(accumulate)
        : "memory");
}
RQ_DEV void tc5_commit(uint64_t* bar) {  // arrives on the mbarrier when every tcgen05.mma issued so far has completed (implies fence::before_thread_sync)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
RQ_DEV void tc5_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
RQ_DEV void tc5_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
RQ_DEV void tc5_ld32(uint32_t taddr, uint32_t (&v)[32]) {  // 32 lanes x 32 consecutive columns: thread = lane (row), v[j] = column j
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(TC5_THREADS, 2) approx_gemm_tc5_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_c,
                                                                         const float* __restrict__ cnorm2, int nq, int K, int D,
                                                                         float* __restrict__ A, int accumulate) {
    extern __shared__ unsigned char tc5_raw[];
    const uint32_t raw_addr = smem_u32(tc5_raw);
    unsigned char* base = tc5_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // swizzle-128B tiles need 1024-byte alignment
    uint64_t* full = reinterpret_cast<uint64_t*>(base + TC5_STAGES * TC5_STAGE_BYTES);  // [STAGES] TMA -> MMA
    uint64_t* empty = full + TC5_STAGES;                                                // [STAGES] MMA -> TMA
    uint64_t* acc_full = empty + TC5_STAGES;                                            // MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.y * TC5_BM, c0 = blockIdx.x * TC5_BN;
    const int nkb = D / TC5_BK;

    if (tid == 0) {
        for (int s = 0; s < TC5_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
    }
    if (warp == 2) {  // one warp allocates the accumulator's 128 TMEM columns (and frees them at the end)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TC5_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % TC5_STAGES;
                const uint32_t ph = (uint32_t)(kb / TC5_STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);  // (passes at once on the stage's first use)
                unsigned char* st = base + (size_t)s * TC5_STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], (uint32_t)TC5_STAGE_BYTES);
                tma_load_2d(st, &tm_y, kb * TC5_BK, q0, &full[s]);
                tma_load_2d(st + TC5_BM * TC5_BK * 4, &tm_c, kb * TC5_BK, c0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % TC5_STAGES;
                const uint32_t ph = (uint32_t)(kb / TC5_STAGES) & 1u;
                mbar_wait(&full[s], ph);
                tc5_fence_after();
                const uint32_t a_addr = smem_u32(base + (size_t)s * TC5_STAGE_BYTES), b_addr = a_addr + TC5_BM * TC5_BK * 4;
#pragma unroll
                for (int k = 0; k < TC5_BK / 8; k++)  // K = 8 TF32 values (32 bytes) per instruction: advance inside the swizzle row
                    tc5_mma(tmem, tc5_smem_desc(a_addr + 32 * k), tc5_smem_desc(b_addr + 32 * k), (kb | k) != 0 ? 1u : 0u);
                tc5_commit(&empty[s]);  // the stage is free once these MMAs have read it
            }
            tc5_commit(acc_full);       // the accumulator tile is complete
        }
    } else {
        mbar_wait(acc_full, 0u);
        tc5_fence_after();
        // A thread owns accumulator row (query) 32 (warp % 4) + lane.  Stored straight from there a warp's instruction would touch
        // 32 different rows, 16 bytes each; instead the warp's 32 x 128 block goes through shared memory (the operand stages are
        // free: every MMA has completed) and leaves as whole rows, 512 contiguous bytes per instruction.
        const int quarter = warp & 3;
        constexpr int TP = TC5_BN + 4;  // pitch in floats: the 8 lanes of a 128-bit store phase hit 32 different banks
        float* tile = reinterpret_cast<float*>(base) + (size_t)quarter * 32 * TP;  // 4 x 16.5 KB <= 3 stages
#pragma unroll 1
        for (int cc = 0; cc < TC5_BN; cc += 32) {
            uint32_t v[32];
            tc5_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cc, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(tile + (size_t)lane * TP + cc + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
        __syncwarp();
        const int c = c0 + 4 * lane;  // this lane's four columns of every row
        if ((K & 3) == 0 && c + 4 <= K) {
            const float4 cn = __ldg(reinterpret_cast<const float4*>(cnorm2 + c));
#pragma unroll 4
            for (int r = 0; r < 32; r++) {
                const int q = q0 + quarter * 32 + r;
                if (q >= nq) break;  // warp-uniform
                const float4 a = *reinterpret_cast<const float4*>(tile + (size_t)r * TP + 4 * lane);
                float4* dst = reinterpret_cast<float4*>(A + (size_t)q * K + c);
                const float4 prev = accumulate ? *dst : cn;
                *dst = make_float4(prev.x - 2.0f * a.x, prev.y - 2.0f * a.y, prev.z - 2.0f * a.z, prev.w - 2.0f * a.w);
            }
        } else {
            for (int r = 0; r < 32; r++) {
                const int q = q0 + quarter * 32 + r;
                if (q >= nq) break;
                for (int j = 0; j < 4; j++)
                    if (c + j < K) {
                        float* dst = A + (size_t)q * K + c + j;
                        const float prev = accumulate ? *dst : __ldg(&cnorm2[c + j]);
                        *dst = prev - 2.0f * tile[(size_t)r * TP + 4 * lane + j];
                    }
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TC5_BN) : "memory");
}

}  // namespace rq
