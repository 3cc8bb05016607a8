// Development probe (sm_100a): the B operand of tcgen05.mma in MN-major form (stored [K][N], N contiguous), as a TMA box
// {64 n, 64 k} with SWIZZLE_128B leaves it in shared memory.  If this works, P.V can read V straight from the qkv planes
// ([keys][head_dim], head_dim contiguous) and the V^T transpose kernel disappears.
//
// One CTA: A[128 x 64] K-major, Bt[64 k x 128 n] loaded as two boxes of 64 n (8 KiB each).  D = A x B for N = 128 with the
// instruction descriptor's b_major bit set, for several (LBO, SBO) candidates of the shared-memory descriptor.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I comprehensive-transformer-tts_b200/csrc \
//             profiles/umma_mn_major_probe.cu -o <out> -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ctts_tc_ptx.cuh"

using namespace ctts;

constexpr int M = 128, KD = 64, NB = 128;
constexpr int NCAND = 6;

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

struct Cand { uint32_t lbo, sbo, kstep; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb,
                                                float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* sa = smem;                    // 128 x 128 B = 16 KiB
    uint8_t* sb = smem + 16 * 1024;        // 2 x (64 k-rows x 128 B) = 16 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 33 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], 16 * 1024 + 16 * 1024);
        tma_load_2d(&ma, &bars[0], sa, 0, 0);
        tma_load_2d(&mb, &bars[0], sb, 0, 0);           // n 0..63
        tma_load_2d(&mb, &bars[0], sb + 8192, 64, 0);   // n 64..127
    }
    mbar_wait(&bars[0], 0);
    // b_major = MN (bit 16)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const Cand cands[NCAND] = {{8192, 1024, 2048}, {1024, 8192, 2048}, {8192, 2048, 2048}, {2048, 8192, 2048}, {8192, 1024, 256}, {1024, 8192, 256}};
    uint32_t phase = 0;
    for (int c = 0; c < NCAND; ++c) {
        if (threadIdx.x == 0) {
            tcgen05_fence_after();
            const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
            for (int k = 0; k < KD / 16; ++k)
                umma_bf16(tmem, desc(a0 + k * 32, 16, 1024), desc(b0 + k * cands[c].kstep, cands[c].lbo, cands[c].sbo), idesc, k ? 1u : 0u);
            umma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], phase);
        phase ^= 1u;
        tcgen05_fence_after();
        float* o = out + ((size_t)c * M + warp * 32 + lane) * NB;
        for (int cc = 0; cc < NB; cc += 16) {
            uint32_t r[16];
            tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + cc, r);
            for (int j = 0; j < 16; ++j) o[cc + j] = __uint_as_float(r[j]);
        }
        tcgen05_fence_before();
        __syncthreads();
    }
    if (warp == 0) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 2;
    EncodeTiledFn enc = (EncodeTiledFn)p;
    std::vector<__nv_bfloat16> ha(M * KD), hb(KD * NB);      // hb = Bt[k][n]
    std::vector<float> fa(M * KD), fb(KD * NB);
    srand(2);
    for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fa[i] = __bfloat162float(ha[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fb[i] = __bfloat162float(hb[i]); }
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, ha.size() * 2);
    cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dout, NCAND * M * NB * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ma, mb;
    cuuint32_t estr[2] = {1, 1};
    {
        cuuint64_t dims[2] = {KD, M}, str[1] = {KD * 2};
        cuuint32_t box[2] = {KD, M};
        if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 3;
    }
    {
        cuuint64_t dims[2] = {NB, KD}, str[1] = {NB * 2};
        cuuint32_t box[2] = {64, KD};
        if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 3;
    }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 36 * 1024);
    probe<<<1, 128, 36 * 1024>>>(ma, mb, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> ho(NCAND * M * NB);
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    const int lbo[NCAND] = {8192, 1024, 8192, 2048, 8192, 1024}, sbo[NCAND] = {1024, 8192, 2048, 8192, 1024, 8192},
              ks[NCAND] = {2048, 2048, 2048, 2048, 256, 256};
    for (int c = 0; c < NCAND; ++c) {
        double worst = 0, worst_lo = 0;
        for (int i = 0; i < M; ++i)
            for (int n = 0; n < NB; ++n) {
                double ref = 0;
                for (int k = 0; k < KD; ++k) ref += (double)fa[i * KD + k] * fb[k * NB + n];
                const double err = fabs(ref - ho[((size_t)c * M + i) * NB + n]);
                worst = fmax(worst, err);
                if (n < 64) worst_lo = fmax(worst_lo, err);
            }
        printf("LBO %5d SBO %5d k-step %4d B : max |err| all 128 n %.3g, first 64 n %.3g  %s\n", lbo[c], sbo[c], ks[c], worst, worst_lo,
               worst < 1e-3 ? "OK" : (worst_lo < 1e-3 ? "first half OK" : "WRONG"));
    }
    return 0;
}
