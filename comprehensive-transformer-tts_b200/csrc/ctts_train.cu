// libctts_b200: kernels of the TRAINING step (training-mode forward pieces and the backward pass), sm_100a.
//
// The reference gets its backward from torch.autograd over ATen ops (train.py:104-123 drives
// model/CompTransTTS.py:64-152 and calls loss.backward()); here every backward step is an explicit kernel reached
// through the C ABI (include/ctts_b200.h, "training step").  The dense contractions of the backward (dgrad = the
// forward engine on re-packed weights, wgrad = ctts_gemm_wgrad) live in ctts_gemm_tc.cu; this file holds
//   * the generic strided FP32 GEMM used for odd shapes (1-, 2-, 4-, 11-wide heads, aligner / attention products in
//     FP32 mode) and as the reference implementation the tensor-core wgrad is tested against,
//   * the HBM-bound pieces: activation / bias backward, LayerNorm backward, BatchNorm (batch statistics) forward and
//     backward, softmax forward / backward over materialised scores, embedding scatter-add, LengthRegulator
//     segment-sum, positional-scale gradient, Philox dropout, weight re-packing for dgrad, operand transposes for
//     wgrad, and the backward of the block-specific kernels (fastformer pooling, GLU, depthwise conv, relative shift,
//     aligner distance attention).
#include "ctts_common.cuh"

#include <math.h>

namespace ctts {

struct TPlanes {
    __nv_bfloat16* p[3];
};

// =====================================================================================================================
// Generic strided FP32 GEMM:  y[z][m,n] = alpha * sum_k A[z][m,k] * B[z][n, k'] (+ y[z][m,n])
//   z = zo * zmod + zi;  element addresses are fully strided (any operand may be "transposed").
//   k is split as (kb, kt) = (k / Kin, k % Kin) so that a reduction can run over (utterance, time); B is read at
//   time kt + shift (zero outside [0, Kin)) with shift = shift0 + z * shift_z -- the tap offset of a conv wgrad.
struct GGAddr {
    int zmod;
    long long a_zo, a_zi, a_m, a_k, a_kb;
    long long b_zo, b_zi, b_n, b_k, b_kb;
    long long y_zo, y_zi, y_m, y_n;
    int Kin, shift0, shift_z;
};

constexpr int XM = 64, XN = 64, XK = 16;

__global__ void __launch_bounds__(256)
gemm_generic_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ Y, int M, int N, int K,
                    const GGAddr g, float alpha, int accumulate) {
    CTTS_PDL_SYNC();
    __shared__ float As[XK][XM + 1];
    __shared__ float Bs[XK][XN + 1];
    const int z = blockIdx.z, zo = z / g.zmod, zi = z - zo * g.zmod;
    const int m0 = blockIdx.x * XM, n0 = blockIdx.y * XN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const float* Az = A + (size_t)zo * g.a_zo + (size_t)zi * g.a_zi;
    const float* Bz = Bm + (size_t)zo * g.b_zo + (size_t)zi * g.b_zi;
    const int shift = g.shift0 + z * g.shift_z;
    const bool a_kfast = g.a_k == 1, b_kfast = g.b_k == 1;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += XK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            {
                const int kk = a_kfast ? (e & 15) : (e >> 6), mm = a_kfast ? (e >> 4) : (e & 63);
                const int k = k0 + kk, m = m0 + mm;
                float v = 0.f;
                if (m < M && k < K) {
                    const int kb = k / g.Kin, kt = k - kb * g.Kin;
                    v = Az[(long long)m * g.a_m + (long long)kb * g.a_kb + (long long)kt * g.a_k];
                }
                As[kk][mm] = v;
            }
            {
                const int kk = b_kfast ? (e & 15) : (e >> 6), nn = b_kfast ? (e >> 4) : (e & 63);
                const int k = k0 + kk, n = n0 + nn;
                float v = 0.f;
                if (n < N && k < K) {
                    const int kb = k / g.Kin, kt = k - kb * g.Kin + shift;
                    if (kt >= 0 && kt < g.Kin) v = Bz[(long long)n * g.b_n + (long long)kb * g.b_kb + (long long)kt * g.b_k];
                }
                Bs[kk][nn] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < XK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* Yz = Y + (size_t)zo * g.y_zo + (size_t)zi * g.y_zi;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float* p = Yz + (long long)m * g.y_m + (long long)n * g.y_n;
            const float v = alpha * acc[i][j];
            *p = accumulate ? (*p + v) : v;
        }
    }
}

// =====================================================================================================================
// Activation backward + bias gradient.
//   forward:  v = (acc + bias) * alpha;  y = act(v)        (ctts_conv1d_gemm / ctts_gemm_split epilogue)
//   backward: dacc[r,n] = dy[r,n] * keep(r) * act'(.) * alpha;   dbias[n] += sum_r dacc[r,n]
//   `ref` is the pre-activation v for GELU / SWISH and the OUTPUT y for RELU / TANH (unused for NONE).
// grid (ceil(N/32), ceil(rows/128), Z): 32 columns x 8 row lanes per CTA; dbias is [Z, N] when Z > 1 (per-utterance
// column sums: the backward of a row broadcast), otherwise [N].
__device__ __forceinline__ float act_grad(float ref, int act) {
    switch (act) {
        case CTTS_ACT_RELU: return ref > 0.f ? 1.f : 0.f;
        case CTTS_ACT_TANH: return 1.f - ref * ref;
        case CTTS_ACT_GELU: {
            const float cdf = 0.5f * (1.f + erff(ref * 0.70710678118654752440f));
            return cdf + ref * 0.3989422804014327f * expf(-0.5f * ref * ref);
        }
        case CTTS_ACT_SWISH: {
            const float s = 1.f / (1.f + expf(-ref));
            return s * (1.f + ref * (1.f - s));
        }
        default: return 1.f;
    }
}

__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ ref, int act, float alpha,
               const int64_t* __restrict__ lens, int T, int rows, int N, float* __restrict__ dz, float* __restrict__ dbias) {
    CTTS_PDL_SYNC();
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + cx;
    const int z = blockIdx.z;
    const size_t zoff = (size_t)z * rows * N;
    const int r0 = blockIdx.y * 128;
    float s = 0.f;
    if (n < N) {
        for (int i = ry; i < 128; i += 8) {
            const int r = r0 + i;
            if (r >= rows) break;
            bool keep = true;
            if (lens) {
                const int gr = z * rows + r;
                const int b = gr / T, t = gr - b * T;
                keep = t < (int)lens[b];
            }
            const size_t o = zoff + (size_t)r * N + n;
            float v = 0.f;
            if (keep) {
                v = dy[o] * alpha;
                if (act != CTTS_ACT_NONE) v *= act_grad(ref[o], act);
            }
            if (dz) dz[o] = v;
            s += v;
        }
    }
    if (dbias) {
        red[ry][cx] = s;
        __syncthreads();
        if (ry == 0 && n < N) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += red[i][cx];
            atomicAdd(dbias + (size_t)z * N + n, t);
        }
    }
}

// The same backward step fused with the operand preparation of the two GEMMs that consume dz: one pass reads dy (+ ref),
// writes dz (fp32, in place allowed), the row-major bf16 planes of dz (A operand of the dgrad GEMM), the TIME-MAJOR planes
// dzT [B, N, Tp] (A operand of the wgrad GEMM) and adds the column sums to dbias.  64 x 64 tiles, one utterance per
// blockIdx.z (rows = time steps of that utterance).
template <int NP>
__global__ void __launch_bounds__(256)
act_bwd_planes_kernel(const float* __restrict__ dy, const float* __restrict__ ref, int act, float alpha,
                      const int64_t* __restrict__ lens, int T, int N, int Tp, float* __restrict__ dz, const TPlanes zp,
                      const TPlanes ztp, float* __restrict__ dbias) {
    CTTS_PDL_SYNC();
    __shared__ float tile[64][65];
    __shared__ float colsum[16][64];
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tid = threadIdx.x;
    const int cq = (tid & 15) * 4, rr = tid >> 4;
    const int len = lens ? (int)lens[b] : T;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = rr + 16 * i, t = r0 + rl, n = n0 + cq;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (t < T && n < N) {      // N % 4 == 0: a float4 never straddles the edge
            const size_t o = ((size_t)b * T + t) * N + n;
            if (t < len) {
                const float4 g = *reinterpret_cast<const float4*>(dy + o);
                v[0] = g.x * alpha; v[1] = g.y * alpha; v[2] = g.z * alpha; v[3] = g.w * alpha;
                if (act != CTTS_ACT_NONE) {
                    const float4 rf = *reinterpret_cast<const float4*>(ref + o);
                    v[0] *= act_grad(rf.x, act); v[1] *= act_grad(rf.y, act);
                    v[2] *= act_grad(rf.z, act); v[3] *= act_grad(rf.w, act);
                }
            }
            if (dz) *reinterpret_cast<float4*>(dz + o) = make_float4(v[0], v[1], v[2], v[3]);
            float rem[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(rem[0], rem[1]);
                const __nv_bfloat162 h23 = __floats2bfloat162_rn(rem[2], rem[3]);
                rem[0] -= __low2float(h01); rem[1] -= __high2float(h01);
                rem[2] -= __low2float(h23); rem[3] -= __high2float(h23);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&h01);
                pk.y = *reinterpret_cast<const uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(zp.p[p] + o) = pk;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tile[rl][cq + j] = v[j];
            cs[j] += v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) colsum[rr][cq + j] = cs[j];
    __syncthreads();
    if (dbias && tid < 64 && n0 + tid < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) t += colsum[i][tid];
        atomicAdd(dbias + n0 + tid, t);
    }
    if (!ztp.p[0]) return;      // the weight gradient reads the row-major planes (ctts_gemm_wgrad_rowmajor): no transposed copy
    const int lane = tid & 31, warp = tid >> 5;
    const int r = r0 + 2 * lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int cl = warp + 8 * i, n = n0 + cl;
        if (n >= N || r >= Tp) continue;
        float a = tile[2 * lane][cl], c = tile[2 * lane + 1][cl];
        const size_t o = ((size_t)b * N + n) * Tp + r;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
            if (r + 1 < Tp) *reinterpret_cast<__nv_bfloat162*>(ztp.p[p] + o) = h;
            else ztp.p[p][o] = __low2bfloat16(h);
            a -= __low2float(h);
            c -= __high2float(h);
        }
    }
}

// y = (res + dropout(x)) * keep: the un-fused residual path when a dropout sits between a GEMM and its residual add
// (transformer_fs2.py:190-192,197-199); the backward of the dropout branch is ctts_dropout on the masked gradient.
__global__ void dropout_add_kernel(const float* __restrict__ x, const float* __restrict__ res, const int64_t* __restrict__ lens,
                                   int T, int C, size_t n, float p, uint64_t seed, uint64_t offset,
                                   const unsigned long long* __restrict__ offset_dev, float* __restrict__ y);

// =====================================================================================================================
// LayerNorm backward (blocks.py:137-156 / nn.LayerNorm).  forward: y = (xhat * gamma + beta) * keep.
//   g = dy * keep;  dxhat = g * gamma;  dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
//   dgamma += sum_r g * xhat;  dbeta += sum_r g.     One warp per row (grid-stride), C <= 1024, C % 4 == 0.
// NV = float4 groups per lane (C <= 128 * NV): the register arrays are sized for the row length at hand, so that C = 256 runs at
// full occupancy instead of carrying the 128 registers a 1024-wide row needs.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ dy, float eps,
                     const int64_t* __restrict__ lens, int rows, int T, int C, float* __restrict__ dx, int accumulate,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
    CTTS_PDL_SYNC();
    extern __shared__ float red[];   // [8][2*C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n4 = C >> 2;
    float4 pg[NV], pb[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) pg[i] = pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
        bool keep = true;
        if (lens) {
            const int b = row / T, t = row - b * T;
            keep = t < (int)lens[b];
        }
        float4* dxr = reinterpret_cast<float4*>(dx + (size_t)row * C);
        if (!keep) {
            if (!accumulate)
                for (int c = lane; c < n4; c += 32) dxr[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
        const float4* gr = reinterpret_cast<const float4*>(dy + (size_t)row * C);
        float4 v[NV], g[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + i * 32;
            if (c < n4) {
                v[i] = xr[c];
                g[i] = gr[c];
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + i * 32;
            if (c < n4) {
                v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
                q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + i * 32;
            if (c < n4) {
                const float4 gm = reinterpret_cast<const float4*>(gamma)[c];
                v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;   // xhat
                pg[i].x += g[i].x * v[i].x; pg[i].y += g[i].y * v[i].y; pg[i].z += g[i].z * v[i].z; pg[i].w += g[i].w * v[i].w;
                pb[i].x += g[i].x; pb[i].y += g[i].y; pb[i].z += g[i].z; pb[i].w += g[i].w;
                g[i].x *= gm.x; g[i].y *= gm.y; g[i].z *= gm.z; g[i].w *= gm.w;   // dxhat
                s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
                s2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
            }
        }
        const float m1 = warp_sum(s1) / (float)C, m2 = warp_sum(s2) / (float)C;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + i * 32;
            if (c < n4) {
                float4 o;
                o.x = rstd * (g[i].x - m1 - v[i].x * m2);
                o.y = rstd * (g[i].y - m1 - v[i].y * m2);
                o.z = rstd * (g[i].z - m1 - v[i].z * m2);
                o.w = rstd * (g[i].w - m1 - v[i].w * m2);
                if (accumulate) {
                    const float4 p = dxr[c];
                    o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                }
                dxr[c] = o;
            }
        }
    }
    // per-CTA reduction of the parameter gradients, then one atomicAdd per column
    float* rg = red + (size_t)warp * 2 * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < n4) {
            reinterpret_cast<float4*>(rg)[c] = pg[i];
            reinterpret_cast<float4*>(rg + C)[c] = pb[i];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += 256) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[(size_t)w * 2 * C + c];
        if (c < C) { if (dgamma) atomicAdd(dgamma + c, t); }
        else if (dbeta) atomicAdd(dbeta + (c - C), t);
    }
}

// =====================================================================================================================
// small elementwise helpers
__global__ void mask_rows_kernel(float* __restrict__ x, const int64_t* __restrict__ lens, int T, int C, size_t total4) {
    CTTS_PDL_SYNC();
    const int c4 = C >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t tok = i / c4;
        const size_t b = tok / T;
        const int t = (int)(tok - b * T);
        if (t >= (int)lens[b]) reinterpret_cast<float4*>(x)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void mask_rows_scalar_kernel(float* __restrict__ x, const int64_t* __restrict__ lens, int T, int C, size_t total) {
    CTTS_PDL_SYNC();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t tok = i / C;
        const size_t b = tok / T;
        const int t = (int)(tok - b * T);
        if (t >= (int)lens[b]) x[i] = 0.f;
    }
}

// y = (accumulate ? y : 0) + a * x
__global__ void axpy_kernel(const float* __restrict__ x, float a, size_t n, int accumulate, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    size_t done = 0;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {      // 16-byte groups, scalar tail
        const size_t n4 = n >> 2;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 v = reinterpret_cast<const float4*>(x)[i];
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (accumulate) o = reinterpret_cast<const float4*>(y)[i];
            reinterpret_cast<float4*>(y)[i] = make_float4(o.x + a * v.x, o.y + a * v.y, o.z + a * v.z, o.w + a * v.w);
        }
        done = n4 << 2;
    }
    for (size_t i = done + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = (accumulate ? y[i] : 0.f) + a * x[i];
}

// y[r, c] = (accumulate ? y : 0) + a * x[r, c] * s[r]
__global__ void rowscale_axpy_kernel(const float* __restrict__ x, const float* __restrict__ s, float a, size_t rows, int C,
                                     int accumulate, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    const size_t total = rows * (size_t)C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        y[i] = (accumulate ? y[i] : 0.f) + a * x[i] * s[i / C];
}

// dtable[idx[r], :] += scale * dy[r, :] * keep(r)      (nn.Embedding backward; rows with idx == skip_idx get nothing:
// padding_idx, blocks.py:10-15).  One warp per row, float atomics.
__global__ void scatter_add_rows_kernel(const float* __restrict__ dy, const int64_t* __restrict__ idx,
                                        const int64_t* __restrict__ lens, int T, int rows, int C, int table_rows,
                                        int skip_idx, float scale, float* __restrict__ dtable) {
    CTTS_PDL_SYNC();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    if (lens) {
        const int b = row / T, t = row - b * T;
        if (t >= (int)lens[b]) return;
    }
    int64_t id = idx[row];
    id = id < 0 ? 0 : (id >= table_rows ? table_rows - 1 : id);
    if ((int)id == skip_idx) return;
    const float* src = dy + (size_t)row * C;
    float* dst = dtable + (size_t)id * C;
    for (int c = lane; c < C; c += 32) atomicAdd(dst + c, scale * src[c]);
}

// LengthRegulator backward: dsrc[b, j, :] (+)= sum_{t in [cum[j-1], min(cum[j], M))} dy[b, t, :]   (segment sum)
__global__ void length_expand_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ cum_lr, int S, int C, int M,
                                         int accumulate, float* __restrict__ dsrc) {
    CTTS_PDL_SYNC();
    const int b = blockIdx.y;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= S) return;
    const int lane = threadIdx.x & 31;
    const int32_t* cum = cum_lr + (size_t)b * S;
    const int t0 = j ? cum[j - 1] : 0;
    int t1 = cum[j];
    if (t1 > M) t1 = M;
    float4* o = reinterpret_cast<float4*>(dsrc + ((size_t)b * S + j) * C);
    for (int c = lane; c < (C >> 2); c += 32) {
        float4 a = accumulate ? o[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = t0; t < t1; ++t) {
            const float4 v = reinterpret_cast<const float4*>(dy + ((size_t)b * M + t) * C)[c];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        o[c] = a;
    }
}

// dalpha += sum_{b, t < len, c} dy[b,t,c] * pe[pos(b,t), c]    (the learnable positional scale, transformer_fs2.py:54-58)
__global__ void add_positions_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ pe,
                                         const int64_t* __restrict__ lens, int T, int C, int pos_mode,
                                         float* __restrict__ dalpha) {
    CTTS_PDL_SYNC();
    __shared__ int s_pos[32];
    __shared__ unsigned s_mask[POS_MAXCH];
    __shared__ float red[8];
    const int b = blockIdx.x;
    const int r0 = blockIdx.y * 32, r1 = min(r0 + 32, T);
    const float* xb = x + (size_t)b * T * C;
    if (pos_mode == 0) {
        block_positions(r0, r1, s_pos, s_mask, [&](int t) { return xb[(size_t)t * C] != 0.f; });
    } else {
        for (int t = r0 + threadIdx.x; t < r1; t += blockDim.x) s_pos[t - r0] = t;
        __syncthreads();
    }
    const int len = lens ? (int)lens[b] : T;
    float s = 0.f;
    for (int i = threadIdx.x; i < (r1 - r0) * C; i += blockDim.x) {
        const int tl = i / C, c = i - tl * C, t = r0 + tl;
        if (t < len) s += dy[((size_t)b * T + t) * C + c] * pe[(size_t)s_pos[tl] * C + c];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        atomicAdd(dalpha, t);
    }
}

// =====================================================================================================================
// Softmax over materialised attention scores (FP32), forward and backward.  S, P, dP: [Z, T, ld]; keys s >= len and query
// rows t >= len give zeros (the key-padding mask of transformer_fs2.py:385-394 / transformer.py:247); lens NULL: no mask.
__global__ void masked_softmax_kernel(const float* __restrict__ S, const int64_t* __restrict__ lens, int H, int T, int Tk,
                                      int ld, size_t rows, int mask_rows, float* __restrict__ P) {
    CTTS_PDL_SYNC();
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const size_t z = row / T;
    const int t = (int)(row - z * T);
    const int len = lens ? min((int)lens[z / H], Tk) : Tk;
    const float* s = S + row * (size_t)ld;
    float* p = P + row * (size_t)ld;
    if (mask_rows && lens && t >= (int)lens[z / H]) {
        for (int j = lane; j < ld; j += 32) p[j] = 0.f;
        return;
    }
    float mx = -INFINITY;
    for (int j = lane; j < len; j += 32) mx = fmaxf(mx, s[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < len; j += 32) sum += expf(s[j] - mx);
    const float inv = 1.f / warp_sum(sum);
    for (int j = lane; j < ld; j += 32) p[j] = (j < len) ? expf(s[j] - mx) * inv : 0.f;
}

// dS = P * (dP - sum_j P * dP) * scale    (in place on dP allowed)
__global__ void softmax_bwd_kernel(const float* __restrict__ P, const float* __restrict__ dP, int Tk, int ld, size_t rows,
                                   float scale, float* __restrict__ dS) {
    CTTS_PDL_SYNC();
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* p = P + row * (size_t)ld;
    const float* g = dP + row * (size_t)ld;
    float* o = dS + row * (size_t)ld;
    float dot = 0.f;
    for (int j = lane; j < Tk; j += 32) dot += p[j] * g[j];
    dot = warp_sum(dot);
    for (int j = lane; j < ld; j += 32) o[j] = (j < Tk) ? p[j] * (g[j] - dot) * scale : 0.f;
}

// =====================================================================================================================
// BatchNorm1d with BATCH statistics over all rows (PostNet modules.py:140-148, conformer conv module conformer.py:465):
// padded frames are part of the statistics, exactly as in the reference.
//   colsum: out[c] += sum_r f(x[r,c])   with f = identity (mean) or (x - mean[c])^2 (variance, second pass)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, const float* __restrict__ center, int rows, int C, float* __restrict__ out) {
    CTTS_PDL_SYNC();
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const int r0 = blockIdx.y * 256;
    float s = 0.f;
    if (c < C) {
        const float mu = center ? center[c] : 0.f;
        for (int i = ry; i < 256; i += 8) {
            const int r = r0 + i;
            if (r >= rows) break;
            const float v = x[(size_t)r * C + c];
            s += center ? (v - mu) * (v - mu) : v;
        }
    }
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][cx];
        atomicAdd(out + c, t);
    }
}

__global__ void scale_vec_kernel(float* __restrict__ v, float a, int n) {
    CTTS_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= a;
}

// y = act(gamma * (x - mean) * rsqrt(var + eps) + beta)  -> fp32 and / or bf16 planes
template <int NP>
__global__ void bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act, size_t rows,
                                  int C, float* __restrict__ y, const TPlanes yp) {
    CTTS_PDL_SYNC();
    const size_t total = rows * (size_t)C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float v = (x[i] - mean[c]) * rsqrtf(var[c] + eps) * gamma[c] + beta[c];
        v = apply_act(v, act);
        if (y) y[i] = v;
        if (NP > 0) {
            float rem = v;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const __nv_bfloat16 h = __float2bfloat16_rn(rem);
                yp.p[p][i] = h;
                rem -= __bfloat162float(h);
            }
        }
    }
}

// y = act(x) -> fp32 and / or planes (training keeps the pre-activation for GELU / Swish, so the activation is its own pass)
template <int NP>
__global__ void act_fwd_kernel(const float* __restrict__ x, size_t n, int act, float* __restrict__ y, const TPlanes yp) {
    CTTS_PDL_SYNC();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = apply_act(x[i], act);
        if (y) y[i] = v;
        if (NP > 0) {
            float rem = v;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const __nv_bfloat16 h = __float2bfloat16_rn(rem);
                yp.p[p][i] = h;
                rem -= __bfloat162float(h);
            }
        }
    }
}

struct CTPlanes {
    const __nv_bfloat16* p[3];
};
// y = sum of the bf16 planes (the fp32 value a tensor-core kernel emitted as planes only)
template <int NP>
__global__ void merge_planes_kernel(const CTPlanes in, size_t n, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int p = NP - 1; p >= 0; --p) v += __bfloat162float(in.p[p][i]);
        y[i] = v;
    }
}

// dst[r, 0:C] = (accumulate ? dst : 0) + src[r, 0:C] with independent row strides (first-row gather / scatter, re-striding)
__global__ void copy_rows_kernel(const float* __restrict__ src, long long src_stride, size_t rows, int C, float* __restrict__ dst,
                                 long long dst_stride, int accumulate) {
    CTTS_PDL_SYNC();
    if ((C & 3) == 0 && (src_stride & 3) == 0 && (dst_stride & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {      // 16-byte groups
        const int c4n = C >> 2;
        const size_t total4 = rows * (size_t)c4n;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
            const size_t r = i / c4n;
            const int c = (int)(i - r * c4n) * 4;
            float4* d = reinterpret_cast<float4*>(dst + r * dst_stride + c);
            float4 v = *reinterpret_cast<const float4*>(src + r * src_stride + c);
            if (accumulate) {
                const float4 o = *d;
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *d = v;
        }
        return;
    }
    const size_t total = rows * (size_t)C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / C;
        const int c = (int)(i - r * C);
        float* d = dst + r * dst_stride + c;
        const float v = src[r * src_stride + c];
        *d = accumulate ? *d + v : v;
    }
}

// running = (1 - momentum) * running + momentum * batch statistic (the variance one unbiased), num_batches_tracked += 1
__global__ void bn_update_running_kernel(const float* __restrict__ mean, const float* __restrict__ var, int rows, float momentum,
                                         int C, float* __restrict__ rmean, float* __restrict__ rvar, int64_t* __restrict__ count) {
    CTTS_PDL_SYNC();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        const float unb = rows > 1 ? (float)rows / (float)(rows - 1) : 1.f;
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean[c];
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * var[c] * unb;
    }
    if (c == 0 && count) count[0] += 1;
}

// BatchNorm backward, pass 1: s1[c] += sum_r g, s2[c] += sum_r g * xhat with g = dy * act'(u), u = gamma * xhat + beta
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     int act, int rows, int C, float* __restrict__ s1, float* __restrict__ s2) {
    CTTS_PDL_SYNC();
    __shared__ float red[2][8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const int r0 = blockIdx.y * 256;
    float a1 = 0.f, a2 = 0.f;
    if (c < C) {
        const float mu = mean[c], rstd = rsqrtf(var[c] + eps), gm = gamma[c], bt = beta[c];
        for (int i = ry; i < 256; i += 8) {
            const int r = r0 + i;
            if (r >= rows) break;
            const size_t o = (size_t)r * C + c;
            const float xh = (x[o] - mu) * rstd;
            float g = dy[o];
            if (act != CTTS_ACT_NONE) {
                const float u = gm * xh + bt;
                g *= act_grad(act == CTTS_ACT_TANH ? tanhf(u) : u, act);
            }
            a1 += g;
            a2 += g * xh;
        }
    }
    red[0][ry][cx] = a1;
    red[1][ry][cx] = a2;
    __syncthreads();
    if (ry == 0 && c < C) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { t1 += red[0][i][cx]; t2 += red[1][i][cx]; }
        atomicAdd(s1 + c, t1);
        atomicAdd(s2 + c, t2);
    }
}

// pass 2: dx = gamma * rstd * (g - s1/n - xhat * s2/n)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    float eps, int act, size_t rows, int C, const float* __restrict__ s1,
                                    const float* __restrict__ s2, float* __restrict__ dx) {
    CTTS_PDL_SYNC();
    const size_t total = rows * (size_t)C;
    const float inv_n = 1.f / (float)rows;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float rstd = rsqrtf(var[c] + eps), gm = gamma[c];
        const float xh = (x[i] - mean[c]) * rstd;
        float g = dy[i];
        if (act != CTTS_ACT_NONE) {
            const float u = gm * xh + beta[c];
            g *= act_grad(act == CTTS_ACT_TANH ? tanhf(u) : u, act);
        }
        dx[i] = gm * rstd * (g - s1[c] * inv_n - xh * s2[c] * inv_n);
    }
}

// =====================================================================================================================
// Dropout: counter-based Philox4x32-10.  Element i of a call draws from counter (i / 4, offset) under key `seed`, so the
// backward pass regenerates the forward mask from (seed, offset) without storing it.  y = x * keep / (1 - p).
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t ctr, uint64_t offset, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = c[i];
}

__global__ void dropout_kernel(const float* __restrict__ x, size_t n, float p, uint64_t seed, uint64_t offset,
                               const unsigned long long* __restrict__ offset_dev, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    if (offset_dev) offset += (uint64_t)offset_dev[0];    // step counter kept on the device: a replayed CUDA graph draws fresh masks
    const float inv = 1.f / (1.f - p);
    const uint32_t thresh = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
    const size_t n4 = (n + 3) >> 2;
    // whole float4 groups move as 16-byte accesses (the tensors of the step are 16-byte aligned and a multiple of 4 long;
    // a ragged tail or a misaligned view takes the scalar path)
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        uint32_t r[4];
        philox4x32_10(seed, (uint64_t)q, offset, r);
        if (vec && q * 4 + 3 < n) {
            const float4 v = reinterpret_cast<const float4*>(x)[q];
            reinterpret_cast<float4*>(y)[q] = make_float4((r[0] >= thresh) ? v.x * inv : 0.f, (r[1] >= thresh) ? v.y * inv : 0.f,
                                                          (r[2] >= thresh) ? v.z * inv : 0.f, (r[3] >= thresh) ? v.w * inv : 0.f);
            continue;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const size_t i = q * 4 + e;
            if (i < n) y[i] = (r[e] >= thresh) ? x[i] * inv : 0.f;
        }
    }
}

__global__ void dropout_add_kernel(const float* __restrict__ x, const float* __restrict__ res, const int64_t* __restrict__ lens,
                                   int T, int C, size_t n, float p, uint64_t seed, uint64_t offset,
                                   const unsigned long long* __restrict__ offset_dev, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    if (offset_dev) offset += (uint64_t)offset_dev[0];
    const float inv = 1.f / (1.f - p);
    const uint32_t thresh = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
    const size_t n4 = (n + 3) >> 2;
    const bool vec4 = (C & 3) == 0 &&
                      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        uint32_t r[4];
        philox4x32_10(seed, (uint64_t)q, offset, r);
        if (vec4 && q * 4 + 3 < n) {      // C % 4 == 0: the four elements belong to one token
            bool keep = true;
            if (lens) {
                const size_t tok = (q * 4) / C;
                const size_t bb = tok / T;
                keep = (int)(tok - bb * T) < (int)lens[bb];
            }
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (keep) {
                const float4 v = reinterpret_cast<const float4*>(x)[q], rs = reinterpret_cast<const float4*>(res)[q];
                o = make_float4(rs.x + ((r[0] >= thresh) ? v.x * inv : 0.f), rs.y + ((r[1] >= thresh) ? v.y * inv : 0.f),
                                rs.z + ((r[2] >= thresh) ? v.z * inv : 0.f), rs.w + ((r[3] >= thresh) ? v.w * inv : 0.f));
            }
            reinterpret_cast<float4*>(y)[q] = o;
            continue;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const size_t i = q * 4 + e;
            if (i >= n) break;
            bool keep = true;
            if (lens) {
                const size_t tok = i / C;
                const size_t bb = tok / T;
                keep = (int)(tok - bb * T) < (int)lens[bb];
            }
            y[i] = keep ? res[i] + ((r[e] >= thresh) ? x[i] * inv : 0.f) : 0.f;
        }
    }
}

// =====================================================================================================================
// Weight re-layout for the backward GEMMs.
//   dgrad:  wd[c, j*N + n] = w[n, c, taps-1-j]   (torch Conv1d / Linear weight [N, Cin, taps]) -- dx = conv(dz, wd)
//   wgrad:  dw[n, c, j] (+)= dwp[n, j*Cin + c]   (packed result of ctts_gemm_wgrad -> torch layout)
__global__ void pack_dgrad_kernel(const float* __restrict__ w, int N, int Cin, int taps, float* __restrict__ wd) {
    CTTS_PDL_SYNC();
    const size_t total = (size_t)N * Cin * taps;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const int j = (int)((i / N) % taps);
        const size_t c = i / ((size_t)N * taps);
        wd[i] = w[((size_t)n * Cin + c) * taps + (taps - 1 - j)];
    }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, int N, int Cin, int taps, int accumulate,
                                    float* __restrict__ dw) {
    CTTS_PDL_SYNC();
    const size_t total = (size_t)N * Cin * taps;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % taps);
        const int c = (int)((i / taps) % Cin);
        const size_t n = i / ((size_t)Cin * taps);
        const float v = dwp[(n * taps + j) * Cin + c];
        dw[i] = accumulate ? dw[i] + v : v;
    }
}

// x fp32 [Z, R, C] (row stride ld_in, column offset c0) -> NP bf16 planes [Z, taps, C, Rp] with
//   out[z, tap, c, r] = x[z, r + tap - taps/2, c]   (zero outside [0, R), and for r >= R)
// i.e. the transposed (rows contiguous) and, for a Conv1d, per-tap pre-shifted operand of the wgrad GEMM, whose reduction
// runs over the rows.  (TMA box origins must be 16-byte aligned in the innermost dimension, so the tap shift cannot be a
// coordinate offset of the load: it is applied here.)
template <int NP>
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ x, int R, int C, int ld_in, int c0, int Rp, int taps, const TPlanes out) {
    CTTS_PDL_SYNC();
    // 64 x 64 tiles: float4 row loads (when the row pitch allows), bf16x2 stores of two neighbouring rows -> every warp
    // store instruction writes one whole 128-byte output row segment per plane
    __shared__ float tile[64][65];
    const int z = blockIdx.z / taps, tap = blockIdx.z - z * taps;
    const int shift = tap - (taps >> 1);
    const int r0 = blockIdx.x * 64, cc0 = blockIdx.y * 64;
    const int tid = threadIdx.x;
    const bool vec = ((ld_in | c0) & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (C & 3) == 0;
    {
        const int cq = (tid & 15) * 4, rr = tid >> 4;      // 16 lanes x float4 cover 64 columns; 16 rows per pass
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = rr + 16 * i, r = r0 + rl, rs = r + shift, c = cc0 + cq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < R && rs >= 0 && rs < R) {
                const float* src = x + ((size_t)z * R + rs) * ld_in + c0 + c;
                if (vec && c + 3 < C) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (c < C) v.x = src[0];
                    if (c + 1 < C) v.y = src[1];
                    if (c + 2 < C) v.z = src[2];
                    if (c + 3 < C) v.w = src[3];
                }
            }
            tile[rl][cq] = v.x; tile[rl][cq + 1] = v.y; tile[rl][cq + 2] = v.z; tile[rl][cq + 3] = v.w;
        }
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    const int r = r0 + 2 * lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int cl = warp + 8 * i, c = cc0 + cl;
        if (c >= C || r >= Rp) continue;
        float a = tile[2 * lane][cl], b = tile[2 * lane + 1][cl];
        const size_t o = (((size_t)z * taps + tap) * C + c) * Rp + r;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            if (r + 1 < Rp) *reinterpret_cast<__nv_bfloat162*>(out.p[p] + o) = h;
            else out.p[p][o] = __low2bfloat16(h);
            a -= __low2float(h);
            b -= __high2float(h);
        }
    }
}

// =====================================================================================================================
// AlignmentEncoder backward, score part (modules.py:1198-1212).  With a = -temp * |q - k|^2, L = log_softmax_S(a) over ALL
// S columns, lp = L + log(prior + 1e-8) (= attn_logprob), soft = softmax over the valid columns of lp:
//   dlp[s] = dlogprob[s] + [s < len] soft[s] * (dsoft[s] - sum_s' soft[s'] dsoft[s'])
//   da[s]  = dlp[s] - exp(L[s]) * sum_s' dlp[s']
// One warp per (b, mel frame).  dq / dk follow from da by two GEMMs (host side).
__global__ void aligner_attention_bwd_kernel(const float* __restrict__ soft, const float* __restrict__ logprob,
                                             const float* __restrict__ prior, const float* __restrict__ dsoft,
                                             const float* __restrict__ dlogprob, const int64_t* __restrict__ src_lens, int M,
                                             int S, size_t rows, float* __restrict__ da) {
    CTTS_PDL_SYNC();
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const size_t b = row / M;
    const int m = (int)(row - b * M);
    const int slen = min((int)src_lens[b], S);
    const float* so = soft + row * (size_t)S;
    const float* lp = logprob + row * (size_t)S;
    float* o = da + row * (size_t)S;
    float dot = 0.f;
    if (dsoft)
        for (int s = lane; s < slen; s += 32) dot += so[s] * dsoft[row * (size_t)S + s];
    dot = warp_sum(dot);
    float tot = 0.f;
    for (int s = lane; s < S; s += 32) {
        float v = dlogprob ? dlogprob[row * (size_t)S + s] : 0.f;
        if (dsoft && s < slen) v += so[s] * (dsoft[row * (size_t)S + s] - dot);
        o[s] = v;
        tot += v;
    }
    tot = warp_sum(tot);
    for (int s = lane; s < S; s += 32) {
        const float L = lp[s] - logf(prior[(b * S + s) * (size_t)M + m] + 1e-8f);
        o[s] -= expf(L) * tot;
    }
}

// =====================================================================================================================
// Block-specific backward kernels.
// GLU (blocks.py:123-134): g = a * sigmoid(gate), h = [a | gate].  dh[r, c] = dg * s;  dh[r, C + c] = dg * a * s * (1 - s)
__global__ void glu_bwd_kernel(const float* __restrict__ h, const float* __restrict__ dg, int C, size_t rows,
                               float* __restrict__ dh) {
    CTTS_PDL_SYNC();
    const size_t total = rows * (size_t)C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / C;
        const int c = (int)(i - r * C);
        const float a = h[r * 2 * C + c], gate = h[r * 2 * C + C + c];
        const float s = 1.f / (1.f + expf(-gate));
        const float g = dg[i];
        dh[r * 2 * C + c] = g * s;
        dh[r * 2 * C + C + c] = g * a * s * (1.f - s);
    }
}

// Depthwise Conv1d ('same', no bias; conformer.py:522-560): y[b,t,c] = sum_j x[b,t+j-K/2,c] * w[c,j]
__global__ void dwconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int K, int T, int C,
                                  float* __restrict__ y) {
    CTTS_PDL_SYNC();
    const int b = blockIdx.z, t = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float* xb = x + (size_t)b * T * C;
    float acc = 0.f;
    const int pad = K >> 1;
    for (int j = 0; j < K; ++j) {
        const int tt = t + j - pad;
        if (tt >= 0 && tt < T) acc = fmaf(xb[(size_t)tt * C + c], w[c * K + j], acc);
    }
    y[((size_t)b * T + t) * C + c] = acc;
}

// dx[b,t,c] = sum_j dy[b, t - j + K/2, c] * w[c, j]
__global__ void dwconv_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, int K, int T, int C,
                                       float* __restrict__ dx) {
    CTTS_PDL_SYNC();
    const int b = blockIdx.z, t = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float* gb = dy + (size_t)b * T * C;
    float acc = 0.f;
    const int pad = K >> 1;
    for (int j = 0; j < K; ++j) {
        const int tt = t - j + pad;
        if (tt >= 0 && tt < T) acc = fmaf(gb[(size_t)tt * C + c], w[c * K + j], acc);
    }
    dx[((size_t)b * T + t) * C + c] = acc;
}

// dw[c, j] += sum_{b,t} dy[b,t,c] * x[b, t+j-K/2, c]; grid (ceil(C/32), K, row chunks of 256 over B*T)
__global__ void __launch_bounds__(256)
dwconv_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x, int K, int B, int T, int C,
                         float* __restrict__ dw) {
    CTTS_PDL_SYNC();
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const int j = blockIdx.y;
    const int r0 = blockIdx.z * 256;
    const int pad = K >> 1;
    float s = 0.f;
    if (c < C) {
        for (int i = ry; i < 256; i += 8) {
            const int r = r0 + i;
            if (r >= B * T) break;
            const int b = r / T, t = r - b * T;
            const int tt = t + j - pad;
            if (tt >= 0 && tt < T) s += dy[(size_t)r * C + c] * x[((size_t)b * T + tt) * C + c];
        }
    }
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][cx];
        atomicAdd(dw + (size_t)c * K + j, t);
    }
}

// Conformer relative shift, backward of the score assembly (conformer.py:405-431):
//   forward  score[i,j] = (content[i,j] + shift(pos)[i,j]) / sqrt_dim
//   backward dcontent = dscore / sqrt_dim;  dpos = unshift(dscore) / sqrt_dim   (positions the shift never reads get 0)
// One thread per (z, i, j) of dpos: pos[i, c] is read by score[i, j] with j = c - (T-1-i) when c >= T-1-i  (j <= i),
// and by score[i-1, j] with j = c + (i-1) + 2 = c + i + 1 when that is < T (row i >= 1 supplies the j > i'+1 part of i' = i-1).
__global__ void relshift_bwd_kernel(const float* __restrict__ dscore, int T, int ld, int ld_out, float inv_sqrt_dim, size_t total,
                                    float* __restrict__ dcontent, float* __restrict__ dpos) {
    CTTS_PDL_SYNC();
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % ld_out);
        const size_t zi = e / ld_out;
        const int i = (int)(zi % T);
        const size_t z = zi / T;
        if (c >= T) {       // padding columns of the output rows
            dcontent[e] = 0.f;
            dpos[e] = 0.f;
            continue;
        }
        const float* ds = dscore + z * (size_t)T * ld;
        dcontent[e] = ds[(size_t)i * ld + c] * inv_sqrt_dim;
        float v = 0.f;
        const int j1 = c - (T - 1 - i);
        if (j1 >= 0) v = ds[(size_t)i * ld + j1];               // j1 <= i automatically (c <= T-1)
        else if (i >= 1) {
            const int j2 = c + i + 1;
            if (j2 < T) v = ds[(size_t)(i - 1) * ld + j2];
        }
        dpos[e] = v * inv_sqrt_dim;
    }
}

// Fastformer additive pooling backward (fastformer.py:308-336).  forward (per b, head h): s_t = logit_t / div + mask_t,
// w = softmax_t(s), pooled[e] = sum_t w_t * val[t,e].
//   dval[t,e] = w_t * dpooled[e];  dw_t = sum_e dpooled[e] * val[t,e];  ds_t = w_t * (dw_t - sum_t' w_t' dw_t');  dlogit = ds / div
// grid (B, ceil(heads/32)), 256 threads = 32 heads x 8 time lanes (as the forward kernel).
__global__ void __launch_bounds__(256)
fastformer_pool_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ values, const int64_t* __restrict__ lens,
                           const float* __restrict__ dpooled, int T, int Hh, int hs, float div, float* __restrict__ dlogits,
                           float* __restrict__ dvalues) {
    CTTS_PDL_SYNC();
    __shared__ float red[8][32];
    const int b = blockIdx.x;
    const int hl = threadIdx.x & 31, tl = threadIdx.x >> 5;
    const int h = blockIdx.y * 32 + hl;
    const bool hv = h < Hh;
    const int len = min((int)lens[b], T);
    const float* lg = logits + (size_t)b * T * Hh;
    const float* vl = values + (size_t)b * T * Hh * hs;
    float dp[4] = {0.f, 0.f, 0.f, 0.f};
    if (hv)
        for (int k = 0; k < hs; ++k) dp[k] = dpooled[(size_t)b * Hh * hs + h * hs + k];
    auto block_reduce = [&](float v, bool is_max) {
        red[tl][hl] = v;
        __syncthreads();
        float r = red[0][hl];
#pragma unroll
        for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i][hl]) : r + red[i][hl];
        __syncthreads();
        return r;
    };
    float mx = -INFINITY;
    if (hv)
        for (int t = tl; t < T; t += 8) mx = fmaxf(mx, lg[(size_t)t * Hh + h] / div + (t < len ? -10000.f : 0.f));
    mx = block_reduce(mx, true);
    float sum = 0.f, dot = 0.f;
    if (hv)
        for (int t = tl; t < T; t += 8) {
            const float e = expf(lg[(size_t)t * Hh + h] / div + (t < len ? -10000.f : 0.f) - mx);
            float dw = 0.f;
            for (int k = 0; k < hs; ++k) dw += dp[k] * vl[((size_t)t * Hh + h) * hs + k];
            sum += e;
            dot += e * dw;
        }
    sum = block_reduce(sum, false);
    dot = block_reduce(dot, false);
    if (!hv) return;
    const float inv = 1.f / sum;
    const float wdot = dot * inv;   // sum_t w_t dw_t
    for (int t = tl; t < T; t += 8) {
        const float w = expf(lg[(size_t)t * Hh + h] / div + (t < len ? -10000.f : 0.f) - mx) * inv;
        float dw = 0.f;
        for (int k = 0; k < hs; ++k) {
            const size_t o = ((size_t)t * Hh + h) * hs + k;
            dw += dp[k] * vl[o];
            dvalues[(size_t)b * T * Hh * hs + o] = w * dp[k];
        }
        dlogits[(size_t)b * T * Hh + (size_t)t * Hh + h] = w * (dw - wdot) / div;
    }
}

// y = a * b (b full or one row per utterance) backward: da = dy * b [masked]; db = dy * a (full) or sum_t dy * a (rowwise)
__global__ void mul_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ bb,
                               int b_rowwise, const int64_t* __restrict__ lens, int T, int C, size_t total, float* __restrict__ da,
                               float* __restrict__ db) {
    CTTS_PDL_SYNC();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t tok = i / C;
        const int c = (int)(i - tok * C);
        const size_t bi = tok / T;
        const int t = (int)(tok - bi * T);
        float g = dy[i];
        if (lens && t >= (int)lens[bi]) g = 0.f;
        const float bv = b_rowwise ? bb[bi * C + c] : bb[i];
        if (da) da[i] = g * bv;
        if (db) {
            if (b_rowwise) atomicAdd(db + bi * C + c, g * a[i]);
            else db[i] = g * a[i];
        }
    }
}

// ---- liu2021 reference encoder (training only; modules.py:332-397, coordconv.py:140-159) --------------------------------
// Activations are channels-last [N, H, W, C] (H = mel frames, W = mel bins).  The 3x3 / stride (1, 2) / pad (1, 1)
// convolutions run as im2col + the dense GEMM engine; BatchNorm2d is BatchNorm over the N*H*W rows.
// AddCoords(rank 2, with_r): channels [x, row coordinate, column coordinate, radius], coordinates scaled to [-1, 1], the
// radius measured from (0.5, 0.5) as the reference does (coordconv.py:36-71).
__global__ void add_coords_kernel(const float* __restrict__ x, int H, int W, size_t total, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const int h = (int)((i / W) % H);
        const float xx = ((float)h / (float)(H - 1)) * 2.f - 1.f;
        const float yy = ((float)w / (float)(W - 1)) * 2.f - 1.f;
        const float rr = sqrtf((xx - 0.5f) * (xx - 0.5f) + (yy - 0.5f) * (yy - 0.5f));
        reinterpret_cast<float4*>(y)[i] = make_float4(x[i], xx, yy, rr);
    }
}

// col[(n, h, wo), (kh*3 + kw)*C + c] = x[n, h + kh - 1, 2*wo + kw - 1, c]   (zero outside)
__global__ void im2col_3x3_s12_kernel(const float* __restrict__ x, int H, int W, int C, int Wo, size_t total,
                                      float* __restrict__ col) {
    CTTS_PDL_SYNC();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int k = (int)((i / C) % 9);
        size_t r = i / ((size_t)9 * C);
        const int wo = (int)(r % Wo); r /= Wo;
        const int h = (int)(r % H);
        const size_t n = r / H;
        const int hh = h + k / 3 - 1, ww = 2 * wo + k % 3 - 1;
        col[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[((n * H + hh) * W + ww) * C + c] : 0.f;
    }
}

// dx[n, h, w, c] = sum over the (kh, kw, wo) that read it of dcol[(n, h - kh + 1, wo), (kh*3 + kw)*C + c]   (gather form)
__global__ void col2im_3x3_s12_kernel(const float* __restrict__ dcol, int H, int W, int C, int Wo, size_t total,
                                      float* __restrict__ dx) {
    CTTS_PDL_SYNC();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        size_t r = i / C;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const size_t n = r / H;
        float acc = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ho = h - kh + 1;
            if (ho < 0 || ho >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int t = w - kw + 1;
                if (t < 0 || (t & 1)) continue;
                const int wo = t >> 1;
                if (wo >= Wo) continue;
                acc += dcol[(((n * H + ho) * Wo + wo) * 9 + (kh * 3 + kw)) * (size_t)C + c];
            }
        }
        dx[i] = acc;
    }
}

// y[r, b, a] = x[r, a, b]
__global__ void permute_last2_kernel(const float* __restrict__ x, int A, int Bd, size_t total, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int a = (int)(i % A);
        const int b = (int)((i / A) % Bd);
        const size_t r = i / ((size_t)A * Bd);
        y[i] = x[(r * A + a) * Bd + b];
    }
}

// GRU backward (single direction, nn.GRU gates r|z|n), one CTA per utterance, BPTT over all T steps.
//   forward per step: gh = W_hh h + b_hh; r = s(gi_r + gh_r); z = s(gi_z + gh_z); n = tanh(gi_n + r * gh_n); h' = (1-z) n + z h
// The hidden states are read from `out` (h_t for every t, saved by the forward); gate pre-activations are recomputed
// per step.  Outputs: dgi [B,T,3H] (gradient w.r.t. x W_ih^T + b_ih) and dgh [B,T,3H] (gradient w.r.t. W_hh h + b_hh;
// differs from dgi in the n gate only).  dW_hh = sum_t dgh_t (x) h_{t-1} and db_hh = sum dgh are GEMM / column-sum calls
// on the host side (no atomics here).
__global__ void gru_bwd_kernel(const float* __restrict__ gi, const float* __restrict__ whh, const float* __restrict__ bhh,
                               const float* __restrict__ out, int out_ld, int out_off, const float* __restrict__ dout,
                               const float* __restrict__ dh_final, int dhf_ld, int T, int H, int reverse,
                               float* __restrict__ dgi, float* __restrict__ dgh_out) {
    CTTS_PDL_SYNC();
    extern __shared__ float sm[];
    const int G = 3 * H;
    float* w = sm;                 // [G][H+1] row-major copy of W_hh (rows = gates)
    float* hprev = w + (size_t)G * (H + 1);   // [H]
    float* gh = hprev + H;         // [G]
    float* dgh = gh + G;           // [G]
    float* dh = dgh + G;           // [H]  gradient flowing into h_t from the future
    const int b = blockIdx.x;
    const int j = threadIdx.x;     // 0 .. G-1
    for (int i = j; i < G * H; i += G) w[(size_t)(i / H) * (H + 1) + (i % H)] = whh[i];
    if (j < H) dh[j] = dh_final ? dh_final[(size_t)b * dhf_ld + j] : 0.f;
    const float* gib = gi + (size_t)b * T * G;
    const float* ob = out + (size_t)b * T * out_ld + out_off;
    const float* dob = dout ? dout + (size_t)b * T * out_ld + out_off : nullptr;
    float* dgib = dgi + (size_t)b * T * G;
    float* dghb = dgh_out + (size_t)b * T * G;
    __syncthreads();
    for (int step = T - 1; step >= 0; --step) {
        const int t = reverse ? (T - 1 - step) : step;
        const int tp = reverse ? t + 1 : t - 1;      // time index of the previous hidden state
        if (j < H) hprev[j] = (step > 0) ? ob[(size_t)tp * out_ld + j] : 0.f;
        __syncthreads();
        float a = bhh[j];
        const float* wr = w + (size_t)j * (H + 1);
        for (int k = 0; k < H; ++k) a = fmaf(wr[k], hprev[k], a);
        gh[j] = a;
        __syncthreads();
        float dh_carry = 0.f;
        if (j < H) {
            const float* g = gib + (size_t)t * G;
            const float r = 1.f / (1.f + expf(-(g[j] + gh[j])));
            const float z = 1.f / (1.f + expf(-(g[H + j] + gh[H + j])));
            const float n = tanhf(g[2 * H + j] + r * gh[2 * H + j]);
            const float dht = dh[j] + (dob ? dob[(size_t)t * out_ld + j] : 0.f);
            const float dn = dht * (1.f - z);
            const float dz = dht * (hprev[j] - n);
            dh_carry = dht * z;
            const float d_n = dn * (1.f - n * n);             // w.r.t. (gi_n + r * gh_n)
            const float d_hn = d_n * r;                       // w.r.t. gh_n
            const float d_r = d_n * gh[2 * H + j] * r * (1.f - r);
            const float d_z = dz * z * (1.f - z);
            dgib[(size_t)t * G + j] = d_r;
            dgib[(size_t)t * G + H + j] = d_z;
            dgib[(size_t)t * G + 2 * H + j] = d_n;
            dghb[(size_t)t * G + j] = d_r;
            dghb[(size_t)t * G + H + j] = d_z;
            dghb[(size_t)t * G + 2 * H + j] = d_hn;
            dgh[j] = d_r;
            dgh[H + j] = d_z;
            dgh[2 * H + j] = d_hn;
        }
        __syncthreads();
        float nd = 0.f;
        if (j < H) {
            nd = dh_carry;      // dh_prev[k] = carry[k] + sum_j dgh[j] * W_hh[j, k]
            for (int jj = 0; jj < G; ++jj) nd = fmaf(dgh[jj], w[(size_t)jj * (H + 1) + j], nd);
        }
        __syncthreads();
        if (j < H) dh[j] = nd;
        __syncthreads();
    }
}

}  // namespace ctts

// =====================================================================================================================
// C ABI
// =====================================================================================================================
using namespace ctts;

static inline int grid_for(size_t n, int per = 256, int cap = 8192) {
    const size_t g = (n + per - 1) / per;
    return (int)(g < (size_t)cap ? (g ? g : 1) : cap);
}

extern "C" {

int ctts_gemm_generic(const float* a, const float* b, float* y, int Z, int zmod, int M, int N, int K, const long long* a_str,
                      const long long* b_str, const long long* y_str, int Kin, int shift0, int shift_z, float alpha,
                      int accumulate, void* stream) {
    CTTS_REQUIRE(a && b && y && a_str && b_str && y_str, "gemm_generic: NULL argument");
    CTTS_REQUIRE(Z > 0 && zmod > 0 && M > 0 && N > 0 && K > 0, "gemm_generic: bad shape Z=%d M=%d N=%d K=%d", Z, M, N, K);
    if (Kin <= 0) Kin = K;
    GGAddr g{zmod, a_str[0], a_str[1], a_str[2], a_str[3], a_str[4], b_str[0], b_str[1], b_str[2], b_str[3], b_str[4],
             y_str[0], y_str[1], y_str[2], y_str[3], Kin, shift0, shift_z};
    dim3 grid((M + XM - 1) / XM, (N + XN - 1) / XN, Z);
    launch_k(gemm_generic_kernel, grid, 256, 0, (cudaStream_t)stream, a, b, y, M, N, K, g, alpha, accumulate);
    return check_launch("gemm_generic");
}

int ctts_act_bwd(const float* dy, const float* ref, int act, float alpha, const int64_t* lens, int Z, int T, int rows, int N,
                 float* dz, float* dbias, void* stream) {
    CTTS_REQUIRE(dy && Z > 0 && rows > 0 && N > 0, "act_bwd: bad arguments");
    CTTS_REQUIRE(act == CTTS_ACT_NONE || ref != nullptr, "act_bwd: activation %d needs its reference tensor", act);
    CTTS_REQUIRE(dz || dbias, "act_bwd: nothing to compute");
    dim3 grid((N + 31) / 32, (rows + 127) / 128, Z);
    launch_k(act_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, dy, ref, act, alpha, lens, T, rows, N, dz, dbias);
    return check_launch("act_bwd");
}

int ctts_act_bwd_planes(const float* dy, const float* ref, int act, float alpha, const int64_t* lens, int B, int T, int N, int Tp,
                        float* dz, int n_planes, void* const* dz_planes, void* const* dzT_planes, float* dbias, void* stream) {
    CTTS_REQUIRE(dy && dz_planes && B > 0 && T > 0 && N > 0 && N % 4 == 0 && Tp >= T && Tp % 2 == 0,
                 "act_bwd_planes: bad arguments (N=%d must be a multiple of 4)", N);
    CTTS_REQUIRE(act == CTTS_ACT_NONE || ref != nullptr, "act_bwd_planes: activation %d needs its reference tensor", act);
    CTTS_REQUIRE(n_planes == 2 || n_planes == 3, "act_bwd_planes: n_planes");
    CTTS_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(ref)) & 15) == 0,
                 "act_bwd_planes: tensors must be 16-byte aligned");
    TPlanes zp{{nullptr, nullptr, nullptr}}, ztp{{nullptr, nullptr, nullptr}};
    for (int p = 0; p < n_planes; ++p) {
        CTTS_REQUIRE(dz_planes[p] && (!dzT_planes || dzT_planes[p]), "act_bwd_planes: NULL plane");
        zp.p[p] = (__nv_bfloat16*)dz_planes[p];
        if (dzT_planes) ztp.p[p] = (__nv_bfloat16*)dzT_planes[p];
    }
    dim3 grid((N + 63) / 64, (Tp + 63) / 64, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_planes == 3) launch_k(act_bwd_planes_kernel<3>, grid, 256, 0, st, dy, ref, act, alpha, lens, T, N, Tp, dz, zp, ztp, dbias);
    else launch_k(act_bwd_planes_kernel<2>, grid, 256, 0, st, dy, ref, act, alpha, lens, T, N, Tp, dz, zp, ztp, dbias);
    return check_launch("act_bwd_planes");
}

int ctts_dropout_add(const float* x, const float* res, const int64_t* lens, int B, int T, int C, float p, unsigned long long seed,
                     unsigned long long offset, const unsigned long long* offset_dev, float* y, void* stream) {
    CTTS_REQUIRE(x && res && y && B > 0 && T > 0 && C > 0 && p >= 0.f && p < 1.f, "dropout_add: bad arguments");
    const size_t n = (size_t)B * T * C;
    launch_k(dropout_add_kernel, grid_for((n + 3) / 4), 256, 0, (cudaStream_t)stream, x, res, lens, T, C, n, p, (uint64_t)seed,
             (uint64_t)offset, offset_dev, y);
    return check_launch("dropout_add");
}

int ctts_layernorm_bwd(const float* x, const float* gamma, const float* dy, float eps, const int64_t* lens, int B, int T,
                       int C, float* dx, int accumulate, float* dgamma, float* dbeta, void* stream) {
    CTTS_REQUIRE(C % 4 == 0 && C <= 1024, "layernorm_bwd: C=%d unsupported", C);
    CTTS_REQUIRE(x && gamma && dy && dx && B > 0 && T > 0, "layernorm_bwd: bad arguments");
    const int rows = B * T;
    int grid = (rows + 7) / 8;
    const size_t sm = (size_t)8 * 2 * C * sizeof(float);
    // grid-stride over the rows: enough CTAs to fill the SMs at the occupancy the row width allows (every CTA ends with
    // 2 C atomics, so not one CTA per 8 rows)
    const int cap = C <= 256 ? 1184 : 592;
    if (grid > cap) grid = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 128) {
        ensure_smem(layernorm_bwd_kernel<1>, sm);
        launch_k(layernorm_bwd_kernel<1>, grid, 256, sm, st, x, gamma, dy, eps, lens, rows, T, C, dx, accumulate, dgamma, dbeta);
    } else if (C <= 256) {
        ensure_smem(layernorm_bwd_kernel<2>, sm);
        launch_k(layernorm_bwd_kernel<2>, grid, 256, sm, st, x, gamma, dy, eps, lens, rows, T, C, dx, accumulate, dgamma, dbeta);
    } else if (C <= 512) {
        ensure_smem(layernorm_bwd_kernel<4>, sm);
        launch_k(layernorm_bwd_kernel<4>, grid, 256, sm, st, x, gamma, dy, eps, lens, rows, T, C, dx, accumulate, dgamma, dbeta);
    } else {
        ensure_smem(layernorm_bwd_kernel<8>, sm);
        launch_k(layernorm_bwd_kernel<8>, grid, 256, sm, st, x, gamma, dy, eps, lens, rows, T, C, dx, accumulate, dgamma, dbeta);
    }
    return check_launch("layernorm_bwd");
}

int ctts_mask_rows(float* x, const int64_t* lens, int B, int T, int C, void* stream) {
    CTTS_REQUIRE(x && lens && C > 0, "mask_rows: bad arguments");
    if (C % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) {
        const size_t total = (size_t)B * T * C;
        launch_k(mask_rows_scalar_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, x, lens, T, C, total);
        return check_launch("mask_rows");
    }
    const size_t total4 = (size_t)B * T * (C / 4);
    launch_k(mask_rows_kernel, grid_for(total4), 256, 0, (cudaStream_t)stream, x, lens, T, C, total4);
    return check_launch("mask_rows");
}

int ctts_axpy(const float* x, float a, size_t n, int accumulate, float* y, void* stream) {
    CTTS_REQUIRE(x && y && n > 0, "axpy: bad arguments");
    launch_k(axpy_kernel, grid_for(n), 256, 0, (cudaStream_t)stream, x, a, n, accumulate, y);
    return check_launch("axpy");
}

int ctts_rowscale_axpy(const float* x, const float* s, float a, int rows, int C, int accumulate, float* y, void* stream) {
    CTTS_REQUIRE(x && s && y && rows > 0 && C > 0, "rowscale_axpy: bad arguments");
    launch_k(rowscale_axpy_kernel, grid_for((size_t)rows * C), 256, 0, (cudaStream_t)stream, x, s, a, (size_t)rows, C, accumulate, y);
    return check_launch("rowscale_axpy");
}

int ctts_scatter_add_rows(const float* dy, const int64_t* idx, const int64_t* lens, int T, int rows, int C, int table_rows,
                          int skip_idx, float scale, float* dtable, void* stream) {
    CTTS_REQUIRE(dy && idx && dtable && rows > 0 && C > 0 && table_rows > 0, "scatter_add_rows: bad arguments");
    launch_k(scatter_add_rows_kernel, (rows + 7) / 8, 256, 0, (cudaStream_t)stream, dy, idx, lens, T > 0 ? T : 1, rows, C,
             table_rows, skip_idx, scale, dtable);
    return check_launch("scatter_add_rows");
}

int ctts_length_expand_bwd(const float* dy, const int32_t* cum_lr, int B, int S, int C, int M, int accumulate, float* dsrc,
                           void* stream) {
    CTTS_REQUIRE(dy && cum_lr && dsrc && C % 4 == 0 && M > 0, "length_expand_bwd: bad arguments");
    dim3 grid((S + 7) / 8, B);
    launch_k(length_expand_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, dy, cum_lr, S, C, M, accumulate, dsrc);
    return check_launch("length_expand_bwd");
}

int ctts_add_positions_bwd(const float* dy, const float* x, const float* pe, int pe_rows, const int64_t* lens, int B, int T,
                           int C, int pos_mode, float* dalpha, void* stream) {
    CTTS_REQUIRE(dy && x && pe && dalpha, "add_positions_bwd: NULL argument");
    CTTS_REQUIRE(pe_rows > T - (pos_mode ? 1 : 0) && T <= 32 * POS_MAXCH, "add_positions_bwd: table / length");
    dim3 grid(B, (T + 31) / 32);
    launch_k(add_positions_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, dy, x, pe, lens, T, C, pos_mode, dalpha);
    return check_launch("add_positions_bwd");
}

int ctts_masked_softmax(const float* S, const int64_t* lens, int H, int Z, int T, int Tk, int ld, int mask_rows, float* P,
                        void* stream) {
    CTTS_REQUIRE(S && P && Z > 0 && T > 0 && Tk > 0 && ld >= Tk && H > 0, "masked_softmax: bad arguments");
    const size_t rows = (size_t)Z * T;
    launch_k(masked_softmax_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, S, lens, H, T, Tk, ld, rows, mask_rows, P);
    return check_launch("masked_softmax");
}

int ctts_softmax_bwd(const float* P, const float* dP, int Z, int T, int Tk, int ld, float scale, float* dS, void* stream) {
    CTTS_REQUIRE(P && dP && dS && Z > 0 && T > 0 && Tk > 0 && ld >= Tk, "softmax_bwd: bad arguments");
    const size_t rows = (size_t)Z * T;
    launch_k(softmax_bwd_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, P, dP, Tk, ld, rows, scale, dS);
    return check_launch("softmax_bwd");
}

int ctts_bn_stats(const float* x, int rows, int C, float* mean, float* var, void* stream) {
    CTTS_REQUIRE(x && mean && var && rows > 0 && C > 0, "bn_stats: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(mean, 0, (size_t)C * sizeof(float), st);
    cudaMemsetAsync(var, 0, (size_t)C * sizeof(float), st);
    dim3 grid((C + 31) / 32, (rows + 255) / 256);
    launch_k(colsum_kernel, grid, 256, 0, st, x, (const float*)nullptr, rows, C, mean);
    launch_k(scale_vec_kernel, (C + 255) / 256, 256, 0, st, mean, 1.f / (float)rows, C);
    launch_k(colsum_kernel, grid, 256, 0, st, x, (const float*)mean, rows, C, var);
    launch_k(scale_vec_kernel, (C + 255) / 256, 256, 0, st, var, 1.f / (float)rows, C);
    return check_launch("bn_stats");
}

int ctts_bn_act_fwd(const float* x, const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                    int act, int rows, int C, float* y, int n_planes, void* const* planes, void* stream) {
    CTTS_REQUIRE(x && mean && var && gamma && beta && rows > 0 && C > 0, "bn_act_fwd: bad arguments");
    CTTS_REQUIRE(n_planes == 0 || n_planes == 2 || n_planes == 3, "bn_act_fwd: n_planes");
    TPlanes tp{{nullptr, nullptr, nullptr}};
    for (int p = 0; p < n_planes; ++p) {
        CTTS_REQUIRE(planes && planes[p], "bn_act_fwd: NULL plane");
        tp.p[p] = (__nv_bfloat16*)planes[p];
    }
    const int grid = grid_for((size_t)rows * C);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_planes == 3) launch_k(bn_act_fwd_kernel<3>, grid, 256, 0, st, x, mean, var, gamma, beta, eps, act, (size_t)rows, C, y, tp);
    else if (n_planes == 2) launch_k(bn_act_fwd_kernel<2>, grid, 256, 0, st, x, mean, var, gamma, beta, eps, act, (size_t)rows, C, y, tp);
    else launch_k(bn_act_fwd_kernel<0>, grid, 256, 0, st, x, mean, var, gamma, beta, eps, act, (size_t)rows, C, y, tp);
    return check_launch("bn_act_fwd");
}

int ctts_act_fwd(const float* x, size_t n, int act, float* y, int n_planes, void* const* planes, void* stream) {
    CTTS_REQUIRE(x && n > 0 && (y || n_planes), "act_fwd: bad arguments");
    CTTS_REQUIRE(n_planes == 0 || n_planes == 2 || n_planes == 3, "act_fwd: n_planes");
    TPlanes tp{{nullptr, nullptr, nullptr}};
    for (int p = 0; p < n_planes; ++p) {
        CTTS_REQUIRE(planes && planes[p], "act_fwd: NULL plane");
        tp.p[p] = (__nv_bfloat16*)planes[p];
    }
    const int grid = grid_for(n);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_planes == 3) launch_k(act_fwd_kernel<3>, grid, 256, 0, st, x, n, act, y, tp);
    else if (n_planes == 2) launch_k(act_fwd_kernel<2>, grid, 256, 0, st, x, n, act, y, tp);
    else launch_k(act_fwd_kernel<0>, grid, 256, 0, st, x, n, act, y, tp);
    return check_launch("act_fwd");
}

int ctts_merge_planes(int n_planes, const void* const* planes, size_t n, float* y, void* stream) {
    CTTS_REQUIRE((n_planes == 2 || n_planes == 3) && planes && y && n > 0, "merge_planes: bad arguments");
    CTPlanes cp{{nullptr, nullptr, nullptr}};
    for (int p = 0; p < n_planes; ++p) {
        CTTS_REQUIRE(planes[p], "merge_planes: NULL plane");
        cp.p[p] = (const __nv_bfloat16*)planes[p];
    }
    if (n_planes == 3) launch_k(merge_planes_kernel<3>, grid_for(n), 256, 0, (cudaStream_t)stream, cp, n, y);
    else launch_k(merge_planes_kernel<2>, grid_for(n), 256, 0, (cudaStream_t)stream, cp, n, y);
    return check_launch("merge_planes");
}

int ctts_copy_rows(const float* src, long long src_stride, int rows, int C, float* dst, long long dst_stride, int accumulate,
                   void* stream) {
    CTTS_REQUIRE(src && dst && rows > 0 && C > 0, "copy_rows: bad arguments");
    launch_k(copy_rows_kernel, grid_for((size_t)rows * C), 256, 0, (cudaStream_t)stream, src, src_stride, (size_t)rows, C, dst,
             dst_stride, accumulate);
    return check_launch("copy_rows");
}

int ctts_bn_update_running(const float* mean, const float* var, int rows, float momentum, int C, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, void* stream) {
    CTTS_REQUIRE(mean && var && running_mean && running_var && C > 0, "bn_update_running: bad arguments");
    launch_k(bn_update_running_kernel, (C + 255) / 256, 256, 0, (cudaStream_t)stream, mean, var, rows, momentum, C, running_mean,
             running_var, num_batches_tracked);
    return check_launch("bn_update_running");
}

int ctts_bn_bwd(const float* dy, const float* x, const float* mean, const float* var, const float* gamma, const float* beta,
                float eps, int act, int rows, int C, float* dx, float* dgamma, float* dbeta, float* workspace, void* stream) {
    CTTS_REQUIRE(dy && x && mean && var && gamma && beta && dx && workspace && rows > 0 && C > 0, "bn_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    float* s1 = workspace;
    float* s2 = workspace + C;
    cudaMemsetAsync(workspace, 0, (size_t)2 * C * sizeof(float), st);
    dim3 grid((C + 31) / 32, (rows + 255) / 256);
    launch_k(bn_bwd_reduce_kernel, grid, 256, 0, st, dy, x, mean, var, gamma, beta, eps, act, rows, C, s1, s2);
    launch_k(bn_bwd_apply_kernel, grid_for((size_t)rows * C), 256, 0, st, dy, x, mean, var, gamma, beta, eps, act, (size_t)rows, C,
             (const float*)s1, (const float*)s2, dx);
    if (dbeta) launch_k(axpy_kernel, grid_for(C), 256, 0, st, (const float*)s1, 1.f, (size_t)C, 1, dbeta);
    if (dgamma) launch_k(axpy_kernel, grid_for(C), 256, 0, st, (const float*)s2, 1.f, (size_t)C, 1, dgamma);
    return check_launch("bn_bwd");
}

int ctts_dropout(const float* x, size_t n, float p, unsigned long long seed, unsigned long long offset,
                 const unsigned long long* offset_dev, float* y, void* stream) {
    CTTS_REQUIRE(x && y && n > 0 && p >= 0.f && p < 1.f, "dropout: bad arguments (p=%f)", (double)p);
    launch_k(dropout_kernel, grid_for((n + 3) / 4), 256, 0, (cudaStream_t)stream, x, n, p, (uint64_t)seed, (uint64_t)offset,
             offset_dev, y);
    return check_launch("dropout");
}

int ctts_pack_conv_weight_dgrad(const float* w, int N, int Cin, int taps, float* wd, void* stream) {
    const size_t total = (size_t)N * Cin * taps;
    CTTS_REQUIRE(w && wd && total > 0, "pack_conv_weight_dgrad: bad arguments");
    launch_k(pack_dgrad_kernel, grid_for(total, 256, 4096), 256, 0, (cudaStream_t)stream, w, N, Cin, taps, wd);
    return check_launch("pack_conv_weight_dgrad");
}

int ctts_unpack_conv_wgrad(const float* dw_packed, int N, int Cin, int taps, int accumulate, float* dw, void* stream) {
    const size_t total = (size_t)N * Cin * taps;
    CTTS_REQUIRE(dw_packed && dw && total > 0, "unpack_conv_wgrad: bad arguments");
    launch_k(unpack_wgrad_kernel, grid_for(total, 256, 4096), 256, 0, (cudaStream_t)stream, dw_packed, N, Cin, taps, accumulate, dw);
    return check_launch("unpack_conv_wgrad");
}

int ctts_split_transpose(const float* x, int Z, int R, int C, int ld_in, int c0, int Rp, int taps, int n_planes,
                         void* const* planes, void* stream) {
    CTTS_REQUIRE(x && planes && Z > 0 && R > 0 && C > 0 && Rp >= R && taps >= 1 && (taps & 1) && (n_planes == 2 || n_planes == 3),
                 "split_transpose: bad arguments");
    TPlanes tp{{nullptr, nullptr, nullptr}};
    for (int p = 0; p < n_planes; ++p) {
        CTTS_REQUIRE(planes[p], "split_transpose: NULL plane");
        tp.p[p] = (__nv_bfloat16*)planes[p];
    }
    CTTS_REQUIRE(Rp % 2 == 0, "split_transpose: Rp must be even");
    dim3 grid((Rp + 63) / 64, (C + 63) / 64, Z * taps);
    if (n_planes == 3) launch_k(split_transpose_kernel<3>, grid, 256, 0, (cudaStream_t)stream, x, R, C, ld_in, c0, Rp, taps, tp);
    else launch_k(split_transpose_kernel<2>, grid, 256, 0, (cudaStream_t)stream, x, R, C, ld_in, c0, Rp, taps, tp);
    return check_launch("split_transpose");
}

int ctts_aligner_attention_bwd(const float* soft, const float* logprob, const float* prior, const float* dsoft,
                               const float* dlogprob, const int64_t* src_lens, int B, int M, int S, float* da, void* stream) {
    CTTS_REQUIRE(soft && logprob && prior && src_lens && da && (dsoft || dlogprob), "aligner_attention_bwd: bad arguments");
    const size_t rows = (size_t)B * M;
    launch_k(aligner_attention_bwd_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, soft, logprob, prior, dsoft,
             dlogprob, src_lens, M, S, rows, da);
    return check_launch("aligner_attention_bwd");
}

int ctts_glu_bwd(const float* h, const float* dg, int rows, int C, float* dh, void* stream) {
    CTTS_REQUIRE(h && dg && dh && rows > 0 && C > 0, "glu_bwd: bad arguments");
    launch_k(glu_bwd_kernel, grid_for((size_t)rows * C), 256, 0, (cudaStream_t)stream, h, dg, C, (size_t)rows, dh);
    return check_launch("glu_bwd");
}

int ctts_dwconv(const float* x, const float* w, int K, int B, int T, int C, float* y, void* stream) {
    CTTS_REQUIRE(x && w && y && B > 0 && T > 0 && C > 0 && (K & 1), "dwconv: bad arguments");
    dim3 grid((C + 127) / 128, T, B);
    launch_k(dwconv_fwd_kernel, grid, 128, 0, (cudaStream_t)stream, x, w, K, T, C, y);
    return check_launch("dwconv");
}

int ctts_dwconv_bwd(const float* dy, const float* x, const float* w, int K, int B, int T, int C, float* dx, float* dw,
                    void* stream) {
    CTTS_REQUIRE(dy && x && w && B > 0 && T > 0 && C > 0 && (K & 1), "dwconv_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dx) {
        dim3 grid((C + 127) / 128, T, B);
        launch_k(dwconv_bwd_data_kernel, grid, 128, 0, st, dy, w, K, T, C, dx);
    }
    if (dw) {
        dim3 grid((C + 31) / 32, K, (B * T + 255) / 256);
        launch_k(dwconv_bwd_weight_kernel, grid, 256, 0, st, dy, x, K, B, T, C, dw);
    }
    return check_launch("dwconv_bwd");
}

int ctts_relshift_bwd(const float* dscore, int Z, int T, int ld, int ld_out, float sqrt_dim, float* dcontent, float* dpos,
                      void* stream) {
    CTTS_REQUIRE(dscore && dcontent && dpos && Z > 0 && T > 0 && ld >= T && ld_out >= T, "relshift_bwd: bad arguments");
    const size_t total = (size_t)Z * T * ld_out;
    launch_k(relshift_bwd_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, dscore, T, ld, ld_out, 1.f / sqrt_dim, total,
             dcontent, dpos);
    return check_launch("relshift_bwd");
}

int ctts_fastformer_pool_bwd(const float* logits, const float* values, const int64_t* lens, const float* dpooled, int B, int T,
                             int heads, int head_size, float* dlogits, float* dvalues, void* stream) {
    CTTS_REQUIRE(logits && values && lens && dpooled && dlogits && dvalues && head_size >= 1 && head_size <= 4,
                 "fastformer_pool_bwd: bad arguments");
    dim3 grid(B, (heads + 31) / 32);
    launch_k(fastformer_pool_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, logits, values, lens, dpooled, T, heads, head_size,
             (float)sqrt((double)head_size), dlogits, dvalues);
    return check_launch("fastformer_pool_bwd");
}

int ctts_mul_bwd(const float* dy, const float* a, const float* b, int b_rowwise, const int64_t* lens, int B, int T, int C,
                 float* da, float* db, void* stream) {
    CTTS_REQUIRE(dy && a && b && (da || db), "mul_bwd: bad arguments");
    const size_t total = (size_t)B * T * C;
    launch_k(mul_bwd_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, dy, a, b, b_rowwise, lens, T, C, total, da, db);
    return check_launch("mul_bwd");
}

int ctts_add_coords(const float* x, int N, int H, int W, float* y, void* stream) {
    CTTS_REQUIRE(x && y && N > 0 && H > 1 && W > 1, "add_coords: bad arguments");
    const size_t total = (size_t)N * H * W;
    launch_k(add_coords_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, x, H, W, total, y);
    return check_launch("add_coords");
}

int ctts_im2col_3x3_s12(const float* x, int N, int H, int W, int C, float* col, void* stream) {
    CTTS_REQUIRE(x && col && N > 0 && H > 0 && W > 0 && C > 0, "im2col_3x3_s12: bad arguments");
    const int Wo = (W + 2 - 3) / 2 + 1;
    const size_t total = (size_t)N * H * Wo * 9 * C;
    launch_k(im2col_3x3_s12_kernel, grid_for(total, 256, 65535), 256, 0, (cudaStream_t)stream, x, H, W, C, Wo, total, col);
    return check_launch("im2col_3x3_s12");
}

int ctts_col2im_3x3_s12(const float* dcol, int N, int H, int W, int C, float* dx, void* stream) {
    CTTS_REQUIRE(dcol && dx && N > 0 && H > 0 && W > 0 && C > 0, "col2im_3x3_s12: bad arguments");
    const int Wo = (W + 2 - 3) / 2 + 1;
    const size_t total = (size_t)N * H * W * C;
    launch_k(col2im_3x3_s12_kernel, grid_for(total, 256, 65535), 256, 0, (cudaStream_t)stream, dcol, H, W, C, Wo, total, dx);
    return check_launch("col2im_3x3_s12");
}

int ctts_permute_last2(const float* x, int rows, int A, int Bd, float* y, void* stream) {
    CTTS_REQUIRE(x && y && rows > 0 && A > 0 && Bd > 0, "permute_last2: bad arguments");
    const size_t total = (size_t)rows * A * Bd;
    launch_k(permute_last2_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, x, A, Bd, total, y);
    return check_launch("permute_last2");
}

int ctts_gru_bwd(const float* gi, const float* w_hh, const float* b_hh, const float* out, int out_ld, int out_off,
                 const float* dout, const float* dh_final, int dhf_ld, int B, int T, int H, int reverse, float* dgi, float* dgh,
                 void* stream) {
    CTTS_REQUIRE(gi && w_hh && b_hh && out && dgi && dgh && (dout || dh_final), "gru_bwd: bad arguments");
    CTTS_REQUIRE(B > 0 && T > 0 && H > 0 && 3 * H <= 1024, "gru_bwd: bad shape B=%d T=%d H=%d", B, T, H);
    const size_t sm = ((size_t)3 * H * (H + 1) + H + 3 * H + 3 * H + H) * sizeof(float);
    CTTS_REQUIRE(sm <= 227 * 1024, "gru_bwd: hidden size %d does not fit shared memory", H);
    ensure_smem(gru_bwd_kernel, sm);
    launch_k(gru_bwd_kernel, B, 3 * H, sm, (cudaStream_t)stream, gi, w_hh, b_hh, out, out_ld, out_off, dout, dh_final, dhf_ld, T, H,
             reverse, dgi, dgh);
    return check_launch("gru_bwd");
}

}  // extern "C"
