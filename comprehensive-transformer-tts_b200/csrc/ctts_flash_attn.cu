// libctts_b200: fused masked self-attention on tcgen05 (head_dim 128, two bf16 operand planes = "bf16x3").
//
// One CTA per (batch*head, 128-query tile).  Scores never leave the SM: S = Q K^T lands in TMEM, the softmax warps turn
// it into probability planes in shared memory, and P V accumulates into a second TMEM region.  No online rescaling of
// the output: the keys are swept TWICE -- pass A computes row maxima (no exponentials), pass B recomputes S, forms
// p = exp(s - m), accumulates the row sums and O += P V; the epilogue divides by the row sum.  Softmax is invariant to
// the shift m, so pass A does not need the exact maximum: it multiplies the HI planes only (one MMA per k-slice instead
// of three, half the K bytes), which bounds p by 2^(a few 1e-2) instead of 1 and changes nothing else.  Pass A is
// latency-bound (one 16 KiB tile per 256 MMA cycles), so its K tiles go through a 4-slot ring laid over the K region.
// The extra sweep costs 1/9 more MMA work than the minimum and saves the 3 x 82 MB round trips (scores, probability
// planes) of the materialised version.
//
//   warp 0      TMA producer: Q once, K hi tiles (pass A, 4 slots), K + V tiles (pass B, 2-stage rings); SWIZZLE_128B
//   warp 1      tcgen05.mma issuer: S (M128 x N64 x K128, 3 plane products) and P V (M128 x N128 x K64)
//   warps 2-9   softmax / epilogue: TMEM lane = query row, two warps per lane quarter split the columns
//
// Replaces transformer_fs2.py:385-394 (F.multi_head_attention_forward) / transformer.py:233-252 on the decoder.
#include "ctts_common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "ctts_tc_ptx.cuh"

namespace ctts {

namespace fa {

constexpr int BQ = 128;      // queries per CTA
constexpr int BKV = 64;      // keys per tile
constexpr int DH = 128;      // head dim
constexpr int Q_TILE = BQ * 64 * 2;       // [128 rows x 64 dims] bf16 = 16 KiB (one plane, one 64-dim k-block)
constexpr int K_TILE = BKV * 64 * 2;      // [64 keys x 64 dims]  = 8 KiB
constexpr int V_TILE = DH * BKV * 2;      // [64 keys x 128 dims] = 16 KiB: two [64 keys x 64 dims] boxes (MN-major B operand)
constexpr int P_TILE = BQ * BKV * 2;      // [128 rows x 64 keys] = 16 KiB
constexpr int Q_BYTES = 4 * Q_TILE;       // 2 planes x 2 k-blocks = 64 KiB
constexpr int K_STAGE = 4 * K_TILE;       // 32 KiB
constexpr int V_STAGE = 2 * V_TILE;       // 32 KiB
constexpr int P_BYTES = 2 * P_TILE;       // 32 KiB
constexpr int KA_SLOT = 2 * K_TILE;       // pass A: hi plane of one key tile (2 k-blocks) = 16 KiB
constexpr int KA_SLOTS = 4;               // ... 4 slots over the 64 KiB K region
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;            // 2 stages
constexpr int OFF_V = OFF_K + 2 * K_STAGE;        // 2 stages
constexpr int OFF_P = OFF_V + 2 * V_STAGE;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int OFF_XCH = OFF_BAR + 256;            // [2][128] float exchange area
constexpr int SMEM_TOTAL = OFF_XCH + 1024 + 1024; // + alignment slack
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

struct Maps {
    CUtensorMap q[2];    // qkv planes [B, T, 3C], box {64, 128, 1}
    CUtensorMap k[2];    // qkv planes,            box {64,  64, 1}
};

__host__ __device__ constexpr uint32_t idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {   // one MUFU op; arguments are <= ~0, huge negatives flush to 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(320, 1)
flash_attention_kernel(const __grid_constant__ Maps tm, const int64_t* __restrict__ lens, int T, int C, int H, float scale_log2e,
                       __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars;            // 1
    uint64_t* k_full = bars + 1;        // [2]
    uint64_t* k_empty = bars + 3;       // [2]
    uint64_t* v_full = bars + 5;        // [2]
    uint64_t* v_empty = bars + 7;       // [2]
    uint64_t* s_full = bars + 9;        // [2]
    uint64_t* s_empty = bars + 11;      // [2]  (4 arrivals: one per softmax warp)
    uint64_t* p_full = bars + 13;       // 4 arrivals
    uint64_t* p_empty = bars + 14;      // 1
    uint64_t* o_full = bars + 15;       // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    uint64_t* ka_full = bars + 17;      // [4]  pass A ring over the K region
    uint64_t* ka_empty = bars + 21;     // [4]

    CTTS_PDL_SYNC();   // lens[] below may come from the previous kernel
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int z = blockIdx.y, b = z / H, h = z - b * H;
    const int q0 = blockIdx.x * BQ;
    const int len = min((int)lens[b], T);
    const int nkv = (len + BKV - 1) / BKV;     // key tiles that hold at least one valid key (len >= 1)

    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            tma_prefetch_desc(&tm.q[p]);
            tma_prefetch_desc(&tm.k[p]);
        }
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 8);      // one arrival per softmax warp
        }
        for (int i = 0; i < KA_SLOTS; ++i) {
            mbar_init(&ka_full[i], 1);
            mbar_init(&ka_empty[i], 1);
        }
        mbar_init(p_full, 8);
        mbar_init(p_empty, 1);
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr uint32_t TMEM_COLS = 256;    // S buffers: columns [0,64) and [64,128); O: columns [128,256)
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // Q: 2 planes x 2 k-blocks of 64 dims
            mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_3d(&tm.q[p], q_full, smem + OFF_Q + (p * 2 + kb) * Q_TILE, h * DH + kb * 64, q0, b);
            // pass A: hi planes of the key tiles, 4-slot ring over the K region
            for (int i = 0; i < nkv; ++i) {
                const int a = i & (KA_SLOTS - 1);
                mbar_wait(&ka_empty[a], ((i >> 2) & 1) ^ 1);
                mbar_expect_tx(&ka_full[a], KA_SLOT);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_3d(&tm.k[0], &ka_full[a], smem + OFF_K + a * KA_SLOT + kb * K_TILE, C + h * DH + kb * 64,
                                i * BKV, b);
            }
            // the pass-B stages reuse the same bytes: every slot's last pass-A MMA must have completed
            for (int a = 0; a < KA_SLOTS && a < nkv; ++a) {
                const int uses = (nkv - a + KA_SLOTS - 1) / KA_SLOTS;
                mbar_wait(&ka_empty[a], (uses - 1) & 1);
            }
            // pass B: both planes of the key tiles and the V tiles, 2-stage rings
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&k_full[s], K_STAGE);
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
                        tma_load_3d(&tm.k[p], &k_full[s], smem + OFF_K + s * K_STAGE + (p * 2 + kb) * K_TILE,
                                    C + h * DH + kb * 64, j * BKV, b);
                mbar_wait(&v_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&v_full[s], V_STAGE);
                // V straight from the qkv planes: [64 keys x 128 dims] per plane as two boxes of 64 dims (the MN-major B
                // operand of P V; no V^T copy in HBM)
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int db = 0; db < 2; ++db)
                        tma_load_3d(&tm.k[p], &v_full[s], smem + OFF_V + s * V_STAGE + p * V_TILE + db * (V_TILE / 2),
                                    2 * C + h * DH + db * 64, j * BKV, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc(BKV), idesc_o = idesc(DH) | UMMA_IDESC_B_MN_MAJOR;
            const uint32_t q_addr = smem_u32(smem + OFF_Q);
            auto issue_a = [&](int i) {   // pass A: S[i & 1] ~ Q_hi K_hi,j^T  (row maxima only)
                const int s = i & 1, a = i & (KA_SLOTS - 1);
                mbar_wait(&ka_full[a], (i >> 2) & 1);
                mbar_wait(&s_empty[s], ((i >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint32_t k_addr = smem_u32(smem + OFF_K + a * KA_SLOT);
                const uint32_t d = tmem_base + (uint32_t)(s * BKV);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t off = k * 32;
                        umma_bf16(d, umma_desc_sw128(q_addr + kb * Q_TILE + off), umma_desc_sw128(k_addr + kb * K_TILE + off),
                                  idesc_s, (kb | k) ? 1u : 0u);
                    }
                umma_commit(&ka_empty[a]);
                umma_commit(&s_full[s]);
            };
            auto issue_s = [&](int j) {   // pass B: S[(nkv + j) & 1] = Q K_j^T, all three plane products
                const int i = nkv + j, s = i & 1, ks = j & 1;
                mbar_wait(&k_full[ks], (j >> 1) & 1);
                mbar_wait(&s_empty[s], ((i >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint32_t k_addr = smem_u32(smem + OFF_K + ks * K_STAGE);
                const uint32_t d = tmem_base + (uint32_t)(s * BKV);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t off = k * 32;
                        const uint64_t qh = umma_desc_sw128(q_addr + (0 * 2 + kb) * Q_TILE + off);
                        const uint64_t ql = umma_desc_sw128(q_addr + (1 * 2 + kb) * Q_TILE + off);
                        const uint64_t kh = umma_desc_sw128(k_addr + (0 * 2 + kb) * K_TILE + off);
                        const uint64_t kl = umma_desc_sw128(k_addr + (1 * 2 + kb) * K_TILE + off);
                        umma_bf16(d, ql, kh, idesc_s, (kb | k) ? 1u : 0u);
                        umma_bf16(d, qh, kl, idesc_s, 1u);
                        umma_bf16(d, qh, kh, idesc_s, 1u);
                    }
                umma_commit(&k_empty[ks]);
                umma_commit(&s_full[s]);
            };
            mbar_wait(q_full, 0);
            for (int i = 0; i < nkv; ++i) issue_a(i);                  // pass A
            issue_s(0);                                                // pass B, software pipelined by one tile
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) issue_s(j + 1);
                const int vs = j & 1;
                mbar_wait(p_full, j & 1);
                mbar_wait(&v_full[vs], (j >> 1) & 1);
                tcgen05_fence_after();
                const uint32_t p_addr = smem_u32(smem + OFF_P);
                const uint32_t v_addr = smem_u32(smem + OFF_V + vs * V_STAGE);
                const uint32_t d = tmem_base + 128u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t off = k * 32;
                    const uint64_t ph = umma_desc_sw128(p_addr + off), pl = umma_desc_sw128(p_addr + P_TILE + off);
                    const uint64_t vh = umma_desc_sw128_mn(v_addr + k * 2048, V_TILE / 2),
                                   vl = umma_desc_sw128_mn(v_addr + V_TILE + k * 2048, V_TILE / 2);
                    umma_bf16(d, pl, vh, idesc_o, (j | k) ? 1u : 0u);
                    umma_bf16(d, ph, vl, idesc_o, 1u);
                    umma_bf16(d, ph, vh, idesc_o, 1u);
                }
                umma_commit(p_empty);
                umma_commit(&v_empty[vs]);
            }
            umma_commit(o_full);
        }
    } else {
        // 8 softmax warps: TMEM lane quarter qq = warp % 4 (32 query rows), column half ch = (warp - 2) / 4 (32 of the 64
        // keys of a tile, 64 of the 128 output dims).  The two warps of a quarter meet twice through shared memory:
        // after pass A (row maximum) and before the epilogue (row sum).
        const int qq = warp & 3;
        const int ch = (warp - 2) >> 2;
        const int row = qq * 32 + lane;                  // query row inside the tile == TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(qq * 32) << 16);
        float* xch = reinterpret_cast<float*>(smem + OFF_XCH);           // [2][128] exchange area (1 KiB)
        // ---- pass A: row maximum only (scores in units of log2: s * scale * log2 e) ----
        float m = -INFINITY;
        for (int i = 0; i < nkv; ++i) {
            const int s = i & 1;
            mbar_wait(&s_full[s], (i >> 1) & 1);
            tcgen05_fence_after();
            uint32_t r[32];
            tmem_ld_32x32(lane_base + (uint32_t)(s * BKV + ch * 32), r);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);
            const int k0 = i * BKV + ch * 32;
            if (k0 + 32 <= len) {                       // warp-uniform: only the last tile is partial
#pragma unroll
                for (int c = 0; c < 32; ++c) m = fmaxf(m, __uint_as_float(r[c]));
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (k0 + c < len) m = fmaxf(m, __uint_as_float(r[c]));
            }
        }
        xch[ch * 128 + row] = m;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qq) : "memory");       // the two warps of this lane quarter
        m = fmaxf(xch[row], xch[128 + row]) * scale_log2e;                  // scale > 0: max commutes with the scaling
        // ---- pass B: unnormalised probabilities -> bf16 planes in shared memory (SWIZZLE_128B A operand of P V) ----
        float l = 0.f;
        uint8_t* p_hi = smem + OFF_P + row * 128;
        uint8_t* p_lo = p_hi + P_TILE;
        for (int j = 0; j < nkv; ++j) {
            const int i = nkv + j, s = i & 1;
            mbar_wait(&s_full[s], (i >> 1) & 1);
            tcgen05_fence_after();
            uint32_t r[32];
            tmem_ld_32x32(lane_base + (uint32_t)(s * BKV + ch * 32), r);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);
            const int k0 = j * BKV + ch * 32;
            // exponentials and the hi/lo split go to registers first: they overlap the previous tile's P V, and only
            // the shared-memory stores wait for it
            const bool full = k0 + 32 <= len;           // warp-uniform: only the last tile is partial
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * e]), scale_log2e, -m));
                float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * e + 1]), scale_log2e, -m));
                if (!full) {
                    if (k0 + 2 * e >= len) p0 = 0.f;
                    if (k0 + 2 * e + 1 >= len) p1 = 0.f;
                }
                l += p0 + p1;
                const __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(p0 - __low2float(hh), p1 - __high2float(hh));
                hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
                lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            mbar_wait(p_empty, (j & 1) ^ 1);          // the previous P V has finished reading the P planes
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {           // my 4 chunks of 8 keys = 16 bytes per plane
                const int phys = (((ch * 4 + c8) ^ (row & 7))) * 16;   // 128-byte swizzle: chunk index XOR (row mod 8)
                *reinterpret_cast<uint4*>(p_hi + phys) = make_uint4(hi[4 * c8], hi[4 * c8 + 1], hi[4 * c8 + 2], hi[4 * c8 + 3]);
                *reinterpret_cast<uint4*>(p_lo + phys) = make_uint4(lo[4 * c8], lo[4 * c8 + 1], lo[4 * c8 + 2], lo[4 * c8 + 3]);
            }
            fence_async_smem();       // make the generic-proxy stores visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qq) : "memory");       // everyone has read the maxima: reuse the area
        xch[ch * 128 + row] = l;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qq) : "memory");
        const float inv_l = 1.f / (xch[row] + xch[128 + row]);
        // ---- epilogue: O (TMEM columns 128..255) / l -> output planes; padded query rows are zero ----
        mbar_wait(o_full, 0);
        tcgen05_fence_after();
        const int t = q0 + row;
        const bool store = t < T;
        const float w = (t < len) ? inv_l : 0.f;
        __nv_bfloat16* oh = out_hi + ((size_t)b * T + (store ? t : 0)) * C + (size_t)h * DH + ch * 64;
        __nv_bfloat16* ol = out_lo + ((size_t)b * T + (store ? t : 0)) * C + (size_t)h * DH + ch * 64;
#pragma unroll 1
        for (int chunk = 0; chunk < 2; ++chunk) {
            uint32_t r[32];
            tmem_ld_32x32(lane_base + 128u + (uint32_t)(ch * 64 + chunk * 32), r);
            if (!store) continue;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float a = __uint_as_float(r[c8 * 8 + e * 2]) * w;
                    const float bb = __uint_as_float(r[c8 * 8 + e * 2 + 1]) * w;
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(a, bb);
                    const __nv_bfloat162 ll = __floats2bfloat162_rn(a - __low2float(hh), bb - __high2float(hh));
                    hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
                    lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                *reinterpret_cast<uint4*>(oh + chunk * 32 + c8 * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(ol + chunk * 32 + c8 * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map3(CUtensorMap* m, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t s1,
                     cuuint64_t s2, cuuint32_t b0, cuuint32_t b1) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return 4; }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t str[2] = {s1 * 2, s2 * 2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, str, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("flash_attention: cuTensorMapEncodeTiled failed: CUresult %d", (int)r); return 4; }
    return 0;
}

}  // namespace fa
}  // namespace ctts

using namespace ctts;

extern "C" int ctts_flash_attention_bf16x3(const void* qkv_hi, const void* qkv_lo, const void* vt_hi, const void* vt_lo,
                                           const int64_t* lens, int B, int T, int C, int H, float scale, void* out_hi,
                                           void* out_lo, void* stream) {
    CTTS_REQUIRE(B > 0 && T > 0 && H > 0 && C == H * fa::DH, "flash_attention: head_dim must be 128 (C=%d, H=%d)", C, H);
    CTTS_REQUIRE(qkv_hi && qkv_lo && lens && out_hi && out_lo, "flash_attention: NULL argument");
    (void)vt_hi; (void)vt_lo;     // V is read from the qkv planes (MN-major operand); the V^T planes are no longer used
    const int Z = B * H;
    const cuuint64_t C3 = (cuuint64_t)3 * C;
    fa::Maps maps;
    const void* q[2] = {qkv_hi, qkv_lo};
    for (int p = 0; p < 2; ++p) {
        if (int e = fa::make_map3(&maps.q[p], q[p], C3, (cuuint64_t)T, (cuuint64_t)B, C3, (cuuint64_t)T * C3, 64, fa::BQ)) return e;
        if (int e = fa::make_map3(&maps.k[p], q[p], C3, (cuuint64_t)T, (cuuint64_t)B, C3, (cuuint64_t)T * C3, 64, fa::BKV)) return e;
    }
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(fa::flash_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fa::SMEM_TOTAL) !=
            cudaSuccess) {
            set_error("flash_attention: cannot reserve %d bytes of shared memory", fa::SMEM_TOTAL);
            return 4;
        }
        configured = true;
    }
    dim3 grid((T + fa::BQ - 1) / fa::BQ, Z);
    launch_k(fa::flash_attention_kernel, grid, 320, fa::SMEM_TOTAL, (cudaStream_t)stream, 
        maps, lens, T, C, H, scale * 1.4426950408889634f, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
    return check_launch("flash_attention");
}
