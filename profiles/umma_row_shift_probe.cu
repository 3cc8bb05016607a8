// Development probe (sm_100a): can a tcgen05 shared-memory descriptor of a K-major SWIZZLE_128B operand start at a row
// that is NOT a multiple of 8 (inside a swizzle atom)?  An implicit-GEMM Conv1d re-reads the same activation rows for
// every tap, shifted by one row; if the descriptor may simply start `s` rows further down, the tile (plus a halo) needs
// to be loaded ONCE per k-block instead of once per tap.
//
// One CTA: TMA-loads A[136 rows x 64] and B[64 x 64] (bf16, K-major, SWIZZLE_128B), then for every shift s = 0..8 and
// both settings of the descriptor's base-offset field (0 and (start >> 7) & 7) computes D = A[s .. s+127] x B^T with four
// K = 16 MMAs and compares with the host result.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I comprehensive-transformer-tts_b200/csrc \
//             profiles/umma_row_shift_probe.cu -o gpurun_out/umma_row_shift_probe -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ctts_tc_ptx.cuh"

using namespace ctts;

constexpr int ROWS = 136, KD = 64, NB = 64;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t base_offset) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(base_offset & 7u) << 49) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb,
                                                float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* sa = smem;                    // 136 rows x 128 B = 17 KiB (round up to 18 KiB)
    uint8_t* sb = smem + 18 * 1024;        // 64 rows x 128 B = 8 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 27 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], ROWS * 128 + NB * 128);
        tma_load_2d(&ma, &bars[0], sa, 0, 0);
        tma_load_2d(&mb, &bars[0], sb, 0, 0);
    }
    mbar_wait(&bars[0], 0);
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t phase = 0;
    for (int mode = 0; mode < 2; ++mode) {
        for (int s = 0; s <= 8; ++s) {
            if (threadIdx.x == 0) {
                tcgen05_fence_after();
                const uint32_t a0 = smem_u32(sa) + s * 128, b0 = smem_u32(sb);
                const uint32_t bo = mode ? ((a0 >> 7) & 7u) : 0u;
                for (int k = 0; k < KD / 16; ++k)
                    umma_bf16(tmem, desc_sw128(a0 + k * 32, bo), desc_sw128(b0 + k * 32, 0), idesc, k ? 1u : 0u);
                umma_commit(&bars[1]);
            }
            mbar_wait(&bars[1], phase);
            phase ^= 1u;
            tcgen05_fence_after();
            float* o = out + ((size_t)(mode * 9 + s) * 128 + warp * 32 + lane) * NB;
            for (int c = 0; c < NB; c += 16) {
                uint32_t r[16];
                tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
                for (int j = 0; j < 16; ++j) o[c + j] = __uint_as_float(r[j]);
            }
            tcgen05_fence_before();
            __syncthreads();
        }
    }
    if (warp == 0) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
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
    std::vector<__nv_bfloat16> ha(256 * KD), hb(NB * KD);
    std::vector<float> fa(256 * KD), fb(NB * KD);
    srand(1);
    for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fa[i] = __bfloat162float(ha[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fb[i] = __bfloat162float(hb[i]); }
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, ha.size() * 2);
    cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dout, 2 * 9 * 128 * NB * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ma, mb;
    cuuint32_t estr[2] = {1, 1};
    {
        cuuint64_t dims[2] = {KD, 256}, str[1] = {KD * 2};
        cuuint32_t box[2] = {KD, ROWS};
        if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 3;
    }
    {
        cuuint64_t dims[2] = {KD, NB}, str[1] = {KD * 2};
        cuuint32_t box[2] = {KD, NB};
        if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 3;
    }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 30 * 1024);
    probe<<<1, 128, 30 * 1024>>>(ma, mb, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> ho(2 * 9 * 128 * NB);
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    for (int mode = 0; mode < 2; ++mode)
        for (int s = 0; s <= 8; ++s) {
            double worst = 0;
            for (int i = 0; i < 128; ++i)
                for (int n = 0; n < NB; ++n) {
                    double ref = 0;
                    for (int k = 0; k < KD; ++k) ref += (double)fa[(i + s) * KD + k] * fb[n * KD + k];
                    worst = fmax(worst, fabs(ref - ho[((size_t)(mode * 9 + s) * 128 + i) * NB + n]));
                }
            printf("base_offset %s  row shift %d : max |err| %.3g  %s\n", mode ? "(addr>>7)&7" : "0          ", s, worst,
                   worst < 1e-3 ? "OK" : "WRONG");
        }
    return 0;
}
