// libctts_b200: FP32 CUDA-core kernels of the CompTransTTS forward path (sm_100a).
//
// These are the kernels that sit UPSTREAM of a quantiser (duration rounding, pitch / energy buckets:
// SURVEY.md section 7 H1) and therefore must be true FP32, plus the HBM-bound integer / indexing work
// (LengthRegulator, embeddings).  The tensor-core engine for the decoder / PostNet lives in
// ctts_gemm_tc.cu.  Reference call sites are cited in include/ctts_b200.h.
#include "ctts_common.cuh"

#include <stdlib.h>

#include <math.h>
#include <string.h>

#include <unordered_map>

namespace ctts {

static thread_local char g_err[512] = "";

void ensure_smem_impl(const void* kernel, size_t bytes) {
    static std::unordered_map<const void*, size_t> granted;
    size_t& g = granted[kernel];
    if (g == 0) g = 48 * 1024;
    if (bytes > g) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        g = bytes;
    }
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("CTTS_PDL");
        return e != nullptr && atoi(e) != 0;
    }();
    return on;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// grid (B, ceil(S / EMB_ROWS))
constexpr int EMB_ROWS = 32;
__global__ void embed_tokens_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ table,
                                    const float* __restrict__ pe, float scale, int S, int C, int vocab,
                                    float* __restrict__ x, float* __restrict__ word,
                                    const int64_t* __restrict__ lens, int pos_mode) {
    CTTS_PDL_SYNC();
    __shared__ int s_pos[EMB_ROWS];
    __shared__ unsigned s_mask[POS_MAXCH];
    const int b = blockIdx.x;
    const int r0 = blockIdx.y * EMB_ROWS, r1 = min(r0 + EMB_ROWS, S);
    const int64_t* tok = tokens + (size_t)b * S;
    if (pos_mode == 0) {
        block_positions(r0, r1, s_pos, s_mask, [&](int t) { return tok[t] != 0; });
    } else {
        for (int t = r0 + threadIdx.x; t < r1; t += blockDim.x) s_pos[t - r0] = t;  // absolute positions (transformer.py:72-74)
        __syncthreads();
    }
    const int c4 = C >> 2;
    const int len = lens ? (int)lens[b] : S;
    for (int i = threadIdx.x; i < (r1 - r0) * c4; i += blockDim.x) {
        const int sl = i / c4, s = r0 + sl, c = (i - sl * c4) << 2;
        int64_t id = tok[s];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        float4 e = *reinterpret_cast<const float4*>(table + (size_t)id * C + c);
        const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)s_pos[sl] * C + c);
        e.x *= scale; e.y *= scale; e.z *= scale; e.w *= scale;
        const size_t o = ((size_t)b * S + s) * C + c;
        *reinterpret_cast<float4*>(word + o) = e;
        *reinterpret_cast<float4*>(x + o) = (s < len) ? make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// grid (B, ceil(T / POS_ROWS)): every CTA re-derives the position count up to its own chunk (T strided loads of
// x[b,t,0] spread over 8 warps) and then updates its rows -- one CTA per utterance left 132 SMs idle, and one warp
// scanning 800 rows serially cost 15 us.
constexpr int POS_ROWS = 32;
__global__ void add_positions_kernel(const float* __restrict__ x, const float* __restrict__ pe,
                                     const float* __restrict__ alpha, const int64_t* __restrict__ lens, int T, int C,
                                     int pos_mode, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    __shared__ int s_pos[POS_ROWS];
    __shared__ unsigned s_mask[POS_MAXCH];
    const int b = blockIdx.x;
    const int r0 = blockIdx.y * POS_ROWS;
    const int r1 = min(r0 + POS_ROWS, T);
    const float* xb = x + (size_t)b * T * C;
    float* yb = y + (size_t)b * T * C;
    if (pos_mode == 0) {
        block_positions(r0, r1, s_pos, s_mask, [&](int t) { return xb[(size_t)t * C] != 0.f; });
    } else {
        for (int t = r0 + threadIdx.x; t < r1; t += blockDim.x) s_pos[t - r0] = t;
        __syncthreads();
    }
    const float a = alpha ? alpha[0] : 1.f;
    const int len = lens ? (int)lens[b] : T;
    const int c4 = C >> 2;
    for (int i = threadIdx.x; i < (r1 - r0) * c4; i += blockDim.x) {
        const int t = r0 + i / c4, c = (i % c4) << 2;
        float4 v = *reinterpret_cast<const float4*>(xb + (size_t)t * C + c);
        if (t < len) {
            const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)s_pos[t - r0] * C + c);
            v.x += a * p.x; v.y += a * p.y; v.z += a * p.z; v.w += a * p.w;
        } else {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        *reinterpret_cast<float4*>(yb + (size_t)t * C + c) = v;
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, two-pass moments in FP32.  C <= 1024, C % 4 == 0.
struct PlanePtrs {
    __nv_bfloat16* p[3];
};

template <int NP>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, const int64_t* __restrict__ lens, int rows,
                                 int T, int C, float* __restrict__ y, const PlanePtrs yp) {
    CTTS_PDL_SYNC();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + (size_t)row * C;
    float4 v[8];
    const int n4 = C >> 2;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + i * 32;
        if (c < n4) {
            v[i] = reinterpret_cast<const float4*>(xr)[c];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + i * 32;
        if (c < n4) {
            const float a = v[i].x - mean, b2 = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b2 * b2) + (c2 * c2 + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    bool keep = true;
    if (lens) {
        const int b = row / T, t = row - b * T;
        keep = t < (int)lens[b];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + i * 32;
        if (c < n4) {
            const float4 g = reinterpret_cast<const float4*>(gamma)[c];
            const float4 bb = reinterpret_cast<const float4*>(beta)[c];
            float4 o;
            o.x = keep ? (v[i].x - mean) * rstd * g.x + bb.x : 0.f;
            o.y = keep ? (v[i].y - mean) * rstd * g.y + bb.y : 0.f;
            o.z = keep ? (v[i].z - mean) * rstd * g.z + bb.z : 0.f;
            o.w = keep ? (v[i].w - mean) * rstd * g.w + bb.w : 0.f;
            if (y) reinterpret_cast<float4*>(y + (size_t)row * C)[c] = o;
            if (NP > 0) {
                float rem[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    __nv_bfloat16 h[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        h[k] = __float2bfloat16_rn(rem[k]);
                        rem[k] -= __bfloat162float(h[k]);
                    }
                    *reinterpret_cast<uint2*>(yp.p[p] + (size_t)row * C + 4 * c) = *reinterpret_cast<uint2*>(h);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM Conv1d / Linear in FP32 on the CUDA cores.
// Tile 128 (t) x 64 (n) x 16 (k); 256 threads, 8x4 outputs per thread.
constexpr int GM = 128, GN = 64, GK = 16, GPAD = 4;

// Addressing of the FP32 GEMM.  z = blockIdx.z is the utterance (conv / linear) or a (batch, head) pair
// (zo = z / mod, zh = z % mod) for the batched attention products of the conformer block.
struct GAddr {
    int mod;
    long long x_so, x_sh; int x_ld;   // A row t:  x + zo*x_so + zh*x_sh + t*x_ld
    long long w_so, w_sh; int w_ld;   // W row n:  w + zo*w_so + zh*w_sh + n*w_ld
    long long y_so, y_sh; int y_ld;   // y row t:  y + zo*y_so + zh*y_sh + t*y_ld   (residual uses the same)
    int lens_div;
};

__global__ void __launch_bounds__(256)
conv1d_gemm_fp32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                        float alpha, const float* __restrict__ col_scale, const float* __restrict__ col_shift, int act,
                        const float* __restrict__ residual, const int64_t* __restrict__ lens, int T, int Cin, int N,
                        int taps, float* __restrict__ y, const GAddr ga) {
    CTTS_PDL_SYNC();
    __shared__ __align__(16) float As[2][GK][GM + GPAD];
    __shared__ __align__(16) float Bs[2][GK][GN + GPAD];
    const int b = blockIdx.z;
    const int zo = b / ga.mod, zh = b - zo * ga.mod;
    const int t0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int pad = taps >> 1;
    const int K = taps * Cin;
    const float* xb = x + (size_t)zo * ga.x_so + (size_t)zh * ga.x_sh;
    const float* wb = w + (size_t)zo * ga.w_so + (size_t)zh * ga.w_sh;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb;
    auto load_gmem = [&](int kk) {
        const int tap = kk / Cin, c0 = kk - tap * Cin;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int l = tid + it * 256;
            const int m = l >> 2, kq = l & 3;
            const int t = t0 + m + tap - pad;
            ra[it] = (t >= 0 && t < T && (t0 + m) < T)
                         ? *reinterpret_cast<const float4*>(xb + (size_t)t * ga.x_ld + c0 + kq * 4)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        {
            const int n = tid >> 2, kq = tid & 3;
            rb = (n0 + n < N) ? *reinterpret_cast<const float4*>(wb + (size_t)(n0 + n) * ga.w_ld + kk + kq * 4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int l = tid + it * 256;
            const int m = l >> 2, kq = l & 3;
            As[buf][kq * 4 + 0][m] = ra[it].x;
            As[buf][kq * 4 + 1][m] = ra[it].y;
            As[buf][kq * 4 + 2][m] = ra[it].z;
            As[buf][kq * 4 + 3][m] = ra[it].w;
        }
        const int n = tid >> 2, kq = tid & 3;
        Bs[buf][kq * 4 + 0][n] = rb.x;
        Bs[buf][kq * 4 + 1][n] = rb.y;
        Bs[buf][kq * 4 + 2][n] = rb.z;
        Bs[buf][kq * 4 + 3][n] = rb.w;
    };

    load_gmem(0);
    store_smem(0);
    __syncthreads();
    int buf = 0;
    for (int kk = 0; kk < K; kk += GK) {
        const bool more = kk + GK < K;
        if (more) load_gmem(kk + GK);
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (more) {
            store_smem(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    const int len = lens ? (int)lens[b / ga.lens_div] : T;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = t0 + ty * 8 + i;
        if (t >= T) continue;
        const size_t row = (size_t)zo * ga.y_so + (size_t)zh * ga.y_sh + (size_t)t * ga.y_ld;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            v *= alpha;
            if (col_scale) v = v * col_scale[n] + col_shift[n];
            v = apply_act(v, act);
            if (residual) v += residual[row + n];
            y[row + n] = (t < len) ? v : 0.f;
        }
    }
}

// Skinny linear layers (a handful of outputs per row, or a handful of rows): the 128 x 64 tile kernel above would run
// 1-16 CTAs with a serial k-loop (20-30 us of pure latency).  One warp per (row, 8 outputs): the row of x stays in
// registers, weight rows stream through coalesced float4 loads, one butterfly reduction per output.
constexpr int SK_NCH = 8;
__global__ void __launch_bounds__(256)
skinny_linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float alpha,
                     const float* __restrict__ col_scale, const float* __restrict__ col_shift, int act,
                     const float* __restrict__ residual, const int64_t* __restrict__ lens, long long rows, int T, int K,
                     int N, int chunks, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gw >= rows * chunks) return;
    const int lane = threadIdx.x & 31;
    const long long row = gw / chunks;
    const int n0 = (int)(gw - row * chunks) * SK_NCH;
    const float* xr = x + (size_t)row * K;
    float acc[SK_NCH];
#pragma unroll
    for (int j = 0; j < SK_NCH; ++j) acc[j] = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
        for (int j = 0; j < SK_NCH; ++j) {
            if (n0 + j < N) {
                const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)(n0 + j) * K + k);
                acc[j] = fmaf(xv.x, wv.x, acc[j]);
                acc[j] = fmaf(xv.y, wv.y, acc[j]);
                acc[j] = fmaf(xv.z, wv.z, acc[j]);
                acc[j] = fmaf(xv.w, wv.w, acc[j]);
            }
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < SK_NCH; ++j) {
        const float sum = warp_sum(acc[j]);
        if (lane == j) mine = sum;
    }
    const int n = n0 + lane;
    if (lane >= SK_NCH || n >= N) return;
    const long long b = row / T;
    const int t = (int)(row - b * T);
    float v = mine;
    if (bias) v += bias[n];
    v *= alpha;
    if (col_scale) v = v * col_scale[n] + col_shift[n];
    v = apply_act(v, act);
    const size_t o = (size_t)row * N + n;
    if (residual) v += residual[o];
    y[o] = (!lens || t < (int)lens[b]) ? v : 0.f;
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int N, int Cin, int taps, float* __restrict__ p) {
    CTTS_PDL_SYNC();
    const size_t total = (size_t)N * Cin * taps;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin);
        const int j = (int)((i / Cin) % taps);
        const size_t n = i / ((size_t)Cin * taps);
        p[i] = w[(n * Cin + c) * taps + j];
    }
}

// ---------------------------------------------------------------------------------------------
// Flash-style masked self-attention, FP32.  32 queries per CTA (128 threads: 4 threads per query row,
// each owning DH/4 output dims), keys streamed in tiles of 32 through shared memory.
template <int DH>
__global__ void __launch_bounds__(128)
attention_fp32_kernel(const float* __restrict__ qkv, const int64_t* __restrict__ lens, int T, int C, float scale,
                      float* __restrict__ out) {
    CTTS_PDL_SYNC();
    constexpr int LD = DH + 4;
    constexpr int DPT = DH / 4;
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;
    float* Ks = Qs + 32 * LD;
    float* Vs = Ks + 32 * LD;
    float* Ps = Vs + 32 * LD;  // [32][33]
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 32;
    const int tid = threadIdx.x, r = tid >> 2, qd = tid & 3;
    const int len = min((int)lens[b], T);
    const size_t ld3 = (size_t)3 * C;
    const float* base = qkv + (size_t)b * T * ld3 + (size_t)h * DH;

    // stage Q (pre-scaled, like the reference: q * head_dim^-0.5 before q k^T)
    for (int i = tid; i < 32 * (DH / 4); i += 128) {
        const int rr = i / (DH / 4), c = (i - rr * (DH / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + rr < T) v = *reinterpret_cast<const float4*>(base + (size_t)(q0 + rr) * ld3 + c);
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        *reinterpret_cast<float4*>(Qs + rr * LD + c) = v;
    }
    float acc[DPT];
#pragma unroll
    for (int i = 0; i < DPT; ++i) acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int k0 = 0; k0 < len; k0 += 32) {
        __syncthreads();
        for (int i = tid; i < 32 * (DH / 4); i += 128) {
            const int rr = i / (DH / 4), c = (i - rr * (DH / 4)) * 4;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + rr < len) {
                const float* p = base + (size_t)(k0 + rr) * ld3 + c;
                kv = *reinterpret_cast<const float4*>(p + C);
                vv = *reinterpret_cast<const float4*>(p + 2 * C);
            }
            *reinterpret_cast<float4*>(Ks + rr * LD + c) = kv;
            *reinterpret_cast<float4*>(Vs + rr * LD + c) = vv;
        }
        __syncthreads();
        float s[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = 0.f;
#pragma unroll 4
        for (int d = 0; d < DH; d += 4) {
            const float4 q4 = *reinterpret_cast<const float4*>(Qs + r * LD + d);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 k4 = *reinterpret_cast<const float4*>(Ks + (qd + 4 * i) * LD + d);
                s[i] = fmaf(q4.x, k4.x, s[i]);
                s[i] = fmaf(q4.y, k4.y, s[i]);
                s[i] = fmaf(q4.z, k4.z, s[i]);
                s[i] = fmaf(q4.w, k4.w, s[i]);
            }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (k0 + qd + 4 * i >= len) s[i] = -INFINITY;
            mx = fmaxf(mx, s[i]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_run, mx);  // finite: every tile holds at least one valid key
        const float corr = expf(m_run - m_new);
        float ps = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float p = expf(s[i] - m_new);
            Ps[r * 33 + qd + 4 * i] = p;
            ps += p;
        }
        ps += __shfl_xor_sync(0xffffffffu, ps, 1);
        ps += __shfl_xor_sync(0xffffffffu, ps, 2);
        l_run = l_run * corr + ps;
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < DPT; ++i) acc[i] *= corr;
        __syncwarp();
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const float p = Ps[r * 33 + k];
#pragma unroll
            for (int i = 0; i < DPT; i += 4) {
                const float4 v4 = *reinterpret_cast<const float4*>(Vs + k * LD + qd * DPT + i);
                acc[i + 0] = fmaf(p, v4.x, acc[i + 0]);
                acc[i + 1] = fmaf(p, v4.y, acc[i + 1]);
                acc[i + 2] = fmaf(p, v4.z, acc[i + 2]);
                acc[i + 3] = fmaf(p, v4.w, acc[i + 3]);
            }
        }
        __syncwarp();
    }
    const int t = q0 + r;
    if (t < T) {
        const bool valid = t < len;
        const float inv = valid ? 1.f / l_run : 0.f;
        float* o = out + ((size_t)b * T + t) * C + (size_t)h * DH + qd * DPT;
#pragma unroll
        for (int i = 0; i < DPT; i += 4)
            *reinterpret_cast<float4*>(o + i) = valid ? make_float4(acc[i] * inv, acc[i + 1] * inv, acc[i + 2] * inv,
                                                                    acc[i + 3] * inv)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void decode_durations_kernel(const float* __restrict__ log_d, float d_control, int n, float* __restrict__ dur) {
    CTTS_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dur[i] = fmaxf(rintf(expf(log_d[i]) - 1.f) * d_control, 0.f);
}

// One warp per utterance: inclusive scans of the LR repeat counts and of the mel2ph (rounded) durations.
__global__ void length_scan_kernel(const float* __restrict__ dur_f, const int64_t* __restrict__ dur_i,
                                   const int64_t* __restrict__ src_lens, int S, int32_t* __restrict__ cum_lr,
                                   int32_t* __restrict__ cum_m2p, int64_t* __restrict__ mel_len,
                                   int64_t* __restrict__ m2p_len) {
    CTTS_PDL_SYNC();
    const int b = blockIdx.x, lane = threadIdx.x;
    const int slen = src_lens ? (int)src_lens[b] : S;
    int run_a = 0, run_b = 0;
    for (int base = 0; base < S; base += 32) {
        const int j = base + lane;
        int ra = 0, rb = 0;
        if (j < S) {
            if (dur_f) {
                const float d = dur_f[(size_t)b * S + j];
                ra = max((int)d, 0);                       // int(expand_size): truncation, modules.py:1241-1242
                rb = (int)rintf(d);                        // torch.round: half-to-even, utils/tools.py:618
            } else {
                const int64_t d = dur_i[(size_t)b * S + j];
                ra = (int)(d > 0 ? d : 0);
                rb = (int)d;
            }
            if (j >= slen) rb = 0;                         // dur * (1 - dur_padding), utils/tools.py:619-620
        }
        int ia = ra, ib = rb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += ua; ib += ub; }
        }
        if (j < S) {
            cum_lr[(size_t)b * S + j] = run_a + ia;
            cum_m2p[(size_t)b * S + j] = run_b + ib;
        }
        run_a += __shfl_sync(0xffffffffu, ia, 31);
        run_b += __shfl_sync(0xffffffffu, ib, 31);
    }
    if (lane == 0) {
        mel_len[b] = run_a;
        if (m2p_len) m2p_len[b] = run_b;
    }
}

__device__ __forceinline__ int upper_bound_i32(const int32_t* __restrict__ a, int n, int v) {
    int lo = 0, hi = n;  // first index with a[idx] > v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] > v) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// One warp per output frame.
__global__ void length_expand_kernel(const float* __restrict__ src, const float* __restrict__ table,
                                     const int64_t* __restrict__ row_index, const int32_t* __restrict__ cum_lr, int S,
                                     int C, int M, int accumulate, float* __restrict__ out,
                                     const int32_t* __restrict__ cum_m2p, int64_t* __restrict__ mel2ph, int M2) {
    CTTS_PDL_SYNC();
    const int b = blockIdx.y;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t < M) {
        const int j = upper_bound_i32(cum_lr + (size_t)b * S, S, t);
        const float* row = nullptr;
        if (j < S) row = table ? table + (size_t)row_index[(size_t)b * S + j] * C : src + ((size_t)b * S + j) * C;
        float4* o = reinterpret_cast<float4*>(out + ((size_t)b * M + t) * C);
        for (int c = lane; c < (C >> 2); c += 32) {
            float4 v = row ? reinterpret_cast<const float4*>(row)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (accumulate) {
                const float4 p = o[c];
                v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
            }
            o[c] = v;
        }
    }
    if (mel2ph && t < M2 && lane == 0) {
        const int j = upper_bound_i32(cum_m2p + (size_t)b * S, S, t);
        mel2ph[(size_t)b * M2 + t] = (j < S) ? (int64_t)(j + 1) : 0;
    }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

__device__ __forceinline__ int64_t f0_to_coarse_dev(float f0) {
    // utils/pitch_tools.py:20-36, all arithmetic in fp32 like the torch branch
    const float mel_min = (float)(1127.0 * 0.06899287148695143);  // 1127*ln(1+50/700)
    const float mel_span = (float)(1127.0 * 0.9444616088408514 - 1127.0 * 0.06899287148695143);  // ln(1+1100/700)
    float m = 1127.f * logf(1.f + f0 / 700.f);
    if (m > 0.f) m = (m - mel_min) * 254.f / mel_span + 1.f;
    if (m <= 1.f) m = 1.f;
    if (m > 255.f) m = 255.f;
    return (int64_t)(m + 0.5f);
}

__global__ void __launch_bounds__(256)
cwt_to_pitch_kernel(const float* __restrict__ cwt, int cwt_stride, const float* __restrict__ scale_w,
                    const float* __restrict__ mean, const float* __restrict__ stdv, int stat_stride, float std_scale,
                    float eps, const float* __restrict__ uv_src, int use_uv, int T, float* __restrict__ f0_norm,
                    float* __restrict__ f0_denorm, int64_t* __restrict__ pitch_idx) {
    CTTS_PDL_SYNC();
    extern __shared__ float rec[];
    __shared__ float red[8];
    const int b = blockIdx.x;
    const float* cb = cwt + (size_t)b * T * cwt_stride;
    float w[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) w[i] = scale_w[i];
    float s = 0.f;
    for (int t = threadIdx.x; t < T; t += 256) {
        const float* c = cb + (size_t)t * cwt_stride;
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 10; ++i) a += c[i] * w[i];
        rec[t] = a;
        s += a;
    }
    const float mu = block_sum_256(s, red) / (float)T;
    float q = 0.f;
    for (int t = threadIdx.x; t < T; t += 256) {
        const float d = rec[t] - mu;
        q += d * d;
    }
    const float sd = sqrtf(block_sum_256(q, red) / (float)(T - 1));  // torch.std: unbiased
    const float m_b = mean[(size_t)b * stat_stride];
    const float s_b = stdv[(size_t)b * stat_stride] * std_scale;
    for (int t = threadIdx.x; t < T; t += 256) {
        const float z = (rec[t] - mu) / sd;
        const float f0 = expf(z * s_b + m_b);
        const float fn = log2f(f0 + eps);
        bool uv = false;
        if (use_uv) uv = uv_src ? (uv_src[(size_t)b * T + t] > 0.f) : (cb[(size_t)t * cwt_stride + 10] > 0.f);
        const float fd = uv ? 0.f : exp2f(fn);
        const size_t o = (size_t)b * T + t;
        if (f0_norm) f0_norm[o] = fn;
        f0_denorm[o] = fd;
        pitch_idx[o] = f0_to_coarse_dev(fd);
    }
}

__global__ void f0_to_pitch_kernel(const float* __restrict__ f0n, const float* __restrict__ uv, int n,
                                   float* __restrict__ f0_denorm, int64_t* __restrict__ pitch_idx) {
    CTTS_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float fd = (uv && uv[i] > 0.f) ? 0.f : exp2f(f0n[i]);
    f0_denorm[i] = fd;
    pitch_idx[i] = f0_to_coarse_dev(fd);
}

// pitch_type 'frame' (modules.py:927-938): f0 (+ voiced / unvoiced logit) per frame from the predictor or the targets
__global__ void frame_pitch_kernel(float* __restrict__ pred, int ldp, const float* __restrict__ f0_target,
                                   const float* __restrict__ uv_target, const int64_t* __restrict__ mel2ph, int use_uv, int n,
                                   float* __restrict__ f0_out, float* __restrict__ f0_denorm, int64_t* __restrict__ pitch_idx) {
    CTTS_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float f0 = f0_target ? f0_target[i] : pred[(size_t)i * ldp];
    bool uv = false;
    if (use_uv) uv = uv_target ? (uv_target[i] > 0.f) : (pred[(size_t)i * ldp + 1] > 0.f);
    const bool pad = mel2ph[i] == 0;
    const float fd = (uv || pad) ? 0.f : exp2f(f0);
    f0_out[i] = pad ? 0.f : f0;          // `f0[pitch_padding] = 0`, in place on the caller's target (modules.py:934-935)
    if (!f0_target && pad) pred[(size_t)i * ldp] = 0.f;   // free-running: f0 IS a view of pitch_pred[:, :, 0] there
    f0_denorm[i] = fd;
    pitch_idx[i] = f0_to_coarse_dev(fd);
}

// pitch_type 'ph': frame index = phoneme bucket gathered through mel2ph (F.pad(pitch, [1, 0]) + torch.gather, modules.py:900-901)
__global__ void gather_index_kernel(const int64_t* __restrict__ idx_ph, const int64_t* __restrict__ mel2ph, int S, int M, int n,
                                    int64_t* __restrict__ out) {
    CTTS_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = i / M;
    const int64_t j = mel2ph[i];
    out[i] = (j > 0 && j <= S) ? idx_ph[(size_t)b * S + (j - 1)] : 0;
}

// get_phoneme_level_pitch (modules.py:874-880, utils/tools.py:47-53): mean of the frame-level f0 over each phoneme's frames
__global__ void phoneme_pitch_kernel(const float* __restrict__ f0, const int64_t* __restrict__ mel2ph,
                                     const int64_t* __restrict__ src_lens, const int64_t* __restrict__ mel_lens, int S, int M,
                                     float* __restrict__ out) {
    CTTS_PDL_SYNC();
    const int b = blockIdx.y;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= S) return;
    const int lane = threadIdx.x & 31;
    const int slen = min((int)src_lens[b], S), mlen = min((int)mel_lens[b], M);
    float sum = 0.f, cnt = 0.f;
    if (j < slen)
        for (int t = lane; t < mlen; t += 32)
            if (mel2ph[(size_t)b * M + t] == j + 1) { sum += f0[(size_t)b * M + t]; cnt += 1.f; }
    sum = warp_sum(sum);
    cnt = warp_sum(cnt);
    if (lane == 0) out[(size_t)b * S + j] = (j < slen) ? sum / fmaxf(cnt, 1.f) : 0.f;
}

__global__ void gather_add_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int rows, int C,
                                  int table_rows, float* __restrict__ x) {
    CTTS_PDL_SYNC();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    int64_t id = idx[row];
    id = id < 0 ? 0 : (id >= table_rows ? table_rows - 1 : id);
    const float4* src = reinterpret_cast<const float4*>(table + (size_t)id * C);
    float4* dst = reinterpret_cast<float4*>(x + (size_t)row * C);
    for (int c = lane; c < (C >> 2); c += 32) {
        float4 v = dst[c];
        const float4 e = src[c];
        v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
        dst[c] = v;
    }
}

__global__ void bucketize_kernel(const float* __restrict__ v, float v_scale, const float* __restrict__ bins, int n_bins,
                                 int n, int64_t* __restrict__ idx) {
    CTTS_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = v[i] * v_scale;
    int lo = 0, hi = n_bins;  // first index with bins[idx] >= x  (torch.bucketize, right=False)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (bins[mid] >= x) hi = mid; else lo = mid + 1;
    }
    idx[i] = lo;
}

__global__ void add_row_broadcast_kernel(const float* __restrict__ x, const float* __restrict__ row, int T, int C,
                                         size_t total4, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    const int c4 = C >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t tok = i / c4;
        const int c = (int)(i - tok * c4);
        const size_t b = tok / T;
        float4 v = reinterpret_cast<const float4*>(x)[i];
        const float4 r = reinterpret_cast<const float4*>(row + b * C)[c];
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        reinterpret_cast<float4*>(y)[i] = v;
    }
}

// 8 elements per thread: two 16-byte loads, one 16-byte store per plane (vec: every pointer 16-byte aligned; the last
// n % 8 elements and misaligned views take the scalar loop)
template <int NP>
__global__ void split_bf16_kernel(const float* __restrict__ x, size_t n, const PlanePtrs out, int vec) {
    CTTS_PDL_SYNC();
    const size_t n8 = vec ? n >> 3 : 0;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n8; g += (size_t)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(x)[2 * g], b = reinterpret_cast<const float4*>(x)[2 * g + 1];
        float rem[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(rem[2 * e], rem[2 * e + 1]);
                rem[2 * e] -= __low2float(h);
                rem[2 * e + 1] -= __high2float(h);
                pk[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            reinterpret_cast<uint4*>(out.p[p])[g] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    for (size_t i = (n8 << 3) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float rem = x[i];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const __nv_bfloat16 h = __float2bfloat16_rn(rem);
            out.p[p][i] = h;
            rem -= __bfloat162float(h);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Fastformer additive-attention pooling (fastformer.py:308-322 and 326-336):
//   s[t,h]  = logits[b,t,h] / sqrt(hs) + (t < len ? -10000 : 0)      (the reference's INVERTED mask, quirk 1)
//   w       = softmax_t(s);   pooled[b, h*hs + e] = sum_t w[t,h] * values[b,t,h*hs+e]
// grid (B, Hh/32), 256 threads = 32 heads x 8 time lanes.
__global__ void __launch_bounds__(256)
fastformer_pool_kernel(const float* __restrict__ logits, const float* __restrict__ values, const int64_t* __restrict__ lens,
                       int T, int Hh, int hs, float div, float* __restrict__ pooled) {
    CTTS_PDL_SYNC();
    __shared__ float red[8][32];
    __shared__ float red2[8][32][4];
    const int b = blockIdx.x;
    const int hl = threadIdx.x & 31, tl = threadIdx.x >> 5;
    const int h = blockIdx.y * 32 + hl;
    const bool hv = h < Hh;
    const int len = min((int)lens[b], T);
    const float* lg = logits + (size_t)b * T * Hh;
    const float* vl = values + (size_t)b * T * Hh * hs;
    float mx = -INFINITY;
    if (hv)
        for (int t = tl; t < T; t += 8) mx = fmaxf(mx, lg[(size_t)t * Hh + h] / div + (t < len ? -10000.f : 0.f));
    red[tl][hl] = mx;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) mx = fmaxf(mx, red[i][hl]);
    __syncthreads();
    float sum = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (hv)
        for (int t = tl; t < T; t += 8) {
            const float e = expf(lg[(size_t)t * Hh + h] / div + (t < len ? -10000.f : 0.f) - mx);
            sum += e;
            for (int k = 0; k < hs; ++k) acc[k] += e * vl[((size_t)t * Hh + h) * hs + k];
        }
    red[tl][hl] = sum;
    for (int k = 0; k < 4; ++k) red2[tl][hl][k] = acc[k];
    __syncthreads();
    if (tl == 0 && hv) {
        float s = 0.f, a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < 8; ++i) {
            s += red[i][hl];
            for (int k = 0; k < 4; ++k) a[k] += red2[i][hl][k];
        }
        for (int k = 0; k < hs; ++k) pooled[(size_t)b * Hh * hs + h * hs + k] = a[k] / s;
    }
}

// y = op(a, b) [masked]: op 0: a + b, op 1: a * b ; b is either full-size or one row per batch element (b_rowwise)
__global__ void binary_kernel(const float* __restrict__ a, const float* __restrict__ bb, int op, int b_rowwise,
                              const int64_t* __restrict__ lens, int T, int C, size_t total4, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    const int c4 = C >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t tok = i / c4;
        const int c = (int)(i - tok * c4);
        const size_t bi = tok / T;
        const int t = (int)(tok - bi * T);
        float4 v = reinterpret_cast<const float4*>(a)[i];
        const float4 r = b_rowwise ? reinterpret_cast<const float4*>(bb + bi * C)[c] : reinterpret_cast<const float4*>(bb)[i];
        if (op == 0) { v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        else { v.x *= r.x; v.y *= r.y; v.z *= r.z; v.w *= r.w; }
        if (lens && t >= (int)lens[bi]) v = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(y)[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Conformer convolution module, elementwise part (conformer.py:459-469):
//   glu:   g[b,t,c] = h[b,t,c] * sigmoid(h[b,t,C+c])                         (GLU over the channel dim, blocks.py:123-134)
//   dwconv: y[b,t,c] = swish( BN( sum_j g[b,t+j-K/2,c] * w[c,j] ) )          (depthwise k=31 'same', eval BatchNorm folded)
__global__ void glu_kernel(const float* __restrict__ h, int C, size_t rows, float* __restrict__ g) {
    CTTS_PDL_SYNC();
    const size_t total = rows * (size_t)C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / C;
        const int c = (int)(i - r * C);
        const float a = h[r * 2 * C + c], gate = h[r * 2 * C + C + c];
        g[i] = a * (1.f / (1.f + expf(-gate)));
    }
}

__global__ void dwconv_bn_swish_kernel(const float* __restrict__ g, const float* __restrict__ w, int K,
                                       const float* __restrict__ scale, const float* __restrict__ shift, int T, int C,
                                       float* __restrict__ y) {
    CTTS_PDL_SYNC();
    const int b = blockIdx.z, t = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float* gb = g + (size_t)b * T * C;
    float acc = 0.f;
    const int pad = K >> 1;
    for (int j = 0; j < K; ++j) {
        const int tt = t + j - pad;
        if (tt >= 0 && tt < T) acc = fmaf(gb[(size_t)tt * C + c], w[c * K + j], acc);
    }
    const float v = acc * scale[c] + shift[c];
    y[((size_t)b * T + t) * C + c] = v / (1.f + expf(-v));
}

// ---------------------------------------------------------------------------------------------
// Conformer relative-position attention, score assembly (conformer.py:405-431):
//   score[z,i,j] = (content[z,i,j] + shift(pos)[z,i,j]) / sqrt(d_model);  P = softmax_j(score)  (NO padding mask, quirk 2)
//   shift(pos)[i,j] = j <= i ? pos[i, T-1-i+j] : (j == i+1 ? 0 : pos[i+1, j-i-2])     (zero-pad + view trick of :423-431)
// One warp per (z, i) row; P is written with row stride ldp (>= T, zero padded) so that it can feed the P.V GEMM.
__global__ void relshift_softmax_kernel(const float* __restrict__ content, const float* __restrict__ pos, int T, int ldp,
                                        float sqrt_dim, size_t rows, float* __restrict__ P) {
    CTTS_PDL_SYNC();
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const size_t z = row / T;
    const int i = (int)(row - z * T);
    const float* cz = content + row * (size_t)T;
    const float* pz = pos + z * (size_t)T * T;
    float* out = P + row * (size_t)ldp;
    auto score = [&](int j) {
        float p;
        if (j <= i) p = pz[(size_t)i * T + (T - 1 - i + j)];
        else if (j == i + 1) p = 0.f;
        else p = pz[(size_t)(i + 1) * T + (j - i - 2)];
        return (cz[j] + p) / sqrt_dim;
    };
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, score(j));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) sum += expf(score(j) - mx);
    const float inv = 1.f / warp_sum(sum);
    for (int j = lane; j < ldp; j += 32) out[j] = (j < T) ? expf(score(j) - mx) * inv : 0.f;
}

// Tensor-core form of the score assembly above: reads content / pos with row stride ld, writes P as bf16 planes
// [Z, T, ldp] (zero padded) -- the A operand of the P.V GEMM -- so the fp32 probabilities never reach HBM.
template <int NP>
__global__ void relshift_softmax_planes_kernel(const float* __restrict__ content, const float* __restrict__ pos, int T, int ld,
                                               int ldp, float sqrt_dim, size_t rows, const PlanePtrs out) {
    CTTS_PDL_SYNC();
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const size_t z = row / T;
    const int i = (int)(row - z * T);
    const float* cz = content + row * (size_t)ld;
    const float* pz = pos + z * (size_t)T * ld;
    auto score = [&](int j) {
        float p;
        if (j <= i) p = pz[(size_t)i * ld + (T - 1 - i + j)];
        else if (j == i + 1) p = 0.f;
        else p = pz[(size_t)(i + 1) * ld + (j - i - 2)];
        return (cz[j] + p) / sqrt_dim;
    };
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, score(j));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) sum += expf(score(j) - mx);
    const float inv = 1.f / warp_sum(sum);
    for (int j = lane; j < ldp; j += 32) {
        float rem = (j < T) ? expf(score(j) - mx) * inv : 0.f;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const __nv_bfloat16 h = __float2bfloat16_rn(rem);
            out.p[p][row * (size_t)ldp + j] = h;
            rem -= __bfloat162float(h);
        }
    }
}

// (x[b, t, c0 + h*DH + d] + bias[h*DH + d]) -> bf16 planes [B*T, H*DHp] with each head zero padded from DH to DHp columns:
// a 32-wide head becomes one 64-wide (SWIZZLE_128B) k-block of the tensor-core GEMMs.
template <int NP>
__global__ void pad_heads_planes_kernel(const float* __restrict__ x, const float* __restrict__ bias, int ld_in, int c0, int H,
                                        int DH, int DHp, size_t total, const PlanePtrs out) {
    CTTS_PDL_SYNC();
    const int Cp = H * DHp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cp = (int)(i % Cp);
        const size_t r = i / Cp;
        const int h = cp / DHp, d = cp - h * DHp;
        float rem = 0.f;
        if (d < DH) {
            rem = x[r * ld_in + c0 + h * DH + d];
            if (bias) rem += bias[h * DH + d];
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const __nv_bfloat16 hh = __float2bfloat16_rn(rem);
            out.p[p][i] = hh;
            rem -= __bfloat162float(hh);
        }
    }
}

// x [B, T, ld_in] (channel offset c0, heads of DH) -> xt [B*H, DH, ldt] (time contiguous, zero padded)
__global__ void transpose_heads_kernel(const float* __restrict__ x, int T, int ld_in, int c0, int H, int DH, int ldt,
                                       float* __restrict__ xt) {
    CTTS_PDL_SYNC();
    __shared__ float tile[32][33];
    const int z = blockIdx.z, b = z / H, h = z % H;
    const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int t = t0 + i, d = d0 + tx;
        tile[i][tx] = (t < T && d < DH) ? x[((size_t)b * T + t) * ld_in + c0 + h * DH + d] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int d = d0 + i, t = t0 + tx;
        if (d < DH && t < ldt) xt[((size_t)z * DH + d) * ldt + t] = tile[tx][i];
    }
}


// ---------------------------------------------------------------------------------------------
// AlignmentEncoder score assembly (modules.py:1198-1212), one warp per (b, mel frame):
//   d[s]     = -temperature * sum_c (q[b,m,c] - k[b,s,c])^2
//   logprob  = log_softmax_s(d) (over ALL S key columns, padded ones included -- quirk 7) + log(prior[b,s,m] + 1e-8)
//   soft     = softmax_s(logprob with -inf at s >= src_len)
// q: [B, M, C], k: [B, S, C], prior: [B, S, M] (the caller's layout, read transposed), outputs [B, M, S].
__global__ void __launch_bounds__(256)
aligner_attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ prior,
                         const int64_t* __restrict__ src_lens, float temperature, int M, int S, int C,
                         float* __restrict__ soft, float* __restrict__ logprob) {
    CTTS_PDL_SYNC();
    extern __shared__ float sm[];
    float* ks = sm;                 // [C][S+1] transposed keys
    float* qs = sm + (size_t)C * (S + 1);   // [8][C]
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * 8 + warp;
    for (int i = threadIdx.x; i < S * C; i += 256) {
        const int s_ = i / C, c = i - s_ * C;
        ks[c * (S + 1) + s_] = k[((size_t)b * S + s_) * C + c];
    }
    if (m < M)
        for (int c = lane; c < C; c += 32) qs[warp * C + c] = q[((size_t)b * M + m) * C + c];
    __syncthreads();
    if (m >= M) return;
    const int slen = min((int)src_lens[b], S);
    const float* qr = qs + warp * C;
    float* lp = logprob + ((size_t)b * M + m) * S;
    float* so = soft + ((size_t)b * M + m) * S;
    // pass 1: distances -> lp (temporarily), running max for the log-softmax
    float mx = -INFINITY;
    for (int s_ = lane; s_ < S; s_ += 32) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c) {
            const float d = qr[c] - ks[c * (S + 1) + s_];
            acc += d * d;
        }
        const float v = -temperature * acc;
        lp[s_] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s_ = lane; s_ < S; s_ += 32) sum += expf(lp[s_] - mx);
    const float lse = mx + logf(warp_sum(sum));
    float mx2 = -INFINITY;
    for (int s_ = lane; s_ < S; s_ += 32) {
        const float v = (lp[s_] - lse) + logf(prior[((size_t)b * S + s_) * M + m] + 1e-8f);
        lp[s_] = v;
        if (s_ < slen) mx2 = fmaxf(mx2, v);
    }
    mx2 = warp_max(mx2);
    float sum2 = 0.f;
    for (int s_ = lane; s_ < slen; s_ += 32) sum2 += expf(lp[s_] - mx2);
    const float inv = 1.f / warp_sum(sum2);
    for (int s_ = lane; s_ < S; s_ += 32) so[s_] = (s_ < slen) ? expf(lp[s_] - mx2) * inv : 0.f;
}

// ---------------------------------------------------------------------------------------------
// Monotonic alignment search (mas_width1 / b_mas, modules.py:36-75), one CTA per utterance.
//   a = log(attn[:M_b, :S_b]); a[0, 1:] = -inf; log_p[i,j] = a[i,j] + max(log_p[i-1,j], log_p[i-1,j-1]) (tie -> j-1, `>=`)
//   backtrack from column S_b-1 of the last row.  prev: workspace uint8 [B, M, S] (1 = came from j-1).
// Outputs: hard [B, M, S] 0/1 (zero outside the valid rectangle), dur [B, S] = column sums (attn_hard.sum(2)).
__global__ void __launch_bounds__(1024)
mas_kernel(const float* __restrict__ attn, const int64_t* __restrict__ src_lens, const int64_t* __restrict__ mel_lens, int M,
           int S, uint8_t* __restrict__ prev, float* __restrict__ hard, float* __restrict__ dur) {
    CTTS_PDL_SYNC();
    extern __shared__ float rowbuf[];  // 2 x S
    const int b = blockIdx.x;
    const int Sb = min((int)src_lens[b], S), Mb = min((int)mel_lens[b], M);
    const float* ab = attn + (size_t)b * M * S;
    uint8_t* pb = prev + (size_t)b * M * S;
    float* hb = hard + (size_t)b * M * S;
    for (size_t i = threadIdx.x; i < (size_t)M * S; i += blockDim.x) hb[i] = 0.f;
    for (int j = threadIdx.x; j < S; j += blockDim.x) dur[(size_t)b * S + j] = 0.f;
    float* cur = rowbuf;
    float* nxt = rowbuf + S;
    for (int j = threadIdx.x; j < Sb; j += blockDim.x) cur[j] = (j == 0) ? logf(ab[0]) : -INFINITY;
    __syncthreads();
    for (int i = 1; i < Mb; ++i) {
        for (int j = threadIdx.x; j < Sb; j += blockDim.x) {
            float best = cur[j];
            uint8_t from_left = 0;
            if (j >= 1 && cur[j - 1] >= cur[j]) { best = cur[j - 1]; from_left = 1; }
            nxt[j] = logf(ab[(size_t)i * S + j]) + best;
            pb[(size_t)i * S + j] = from_left;
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
    }
    if (threadIdx.x == 0 && Mb > 0 && Sb > 0) {
        float* db = dur + (size_t)b * S;
        int j = Sb - 1;
        for (int i = Mb - 1; i >= 0; --i) {
            hb[(size_t)i * S + j] = 1.f;
            db[j] += 1.f;
            if (i > 0) j -= pb[(size_t)i * S + j];
        }
        // `opt[0, curr_text_idx] = 1` after the loop with prev_ind[0, :] == 0 -> column 0 (modules.py:63)
        if (hb[0] != 1.f) { hb[0] = 1.f; db[0] += 1.f; }
    }
}

// Phoneme-level averaging of a frame-level feature by hard durations, IN PLACE and sequentially like the reference
// (utils/tools.py:56-66 via modules.py:882-888): out[b, i] = mean(e[pos:pos+d_i]) if d_i > 0 else 0, pos += d_i,
// where `e` is the frame array being overwritten as it goes.  One thread per utterance (S <= a few hundred).
__global__ void phoneme_energy_kernel(const float* __restrict__ dur, const int64_t* __restrict__ src_lens,
                                      const float* __restrict__ energy, int B, int S, int M, float* __restrict__ work,
                                      float* __restrict__ out) {
    CTTS_PDL_SYNC();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float* e = work + (size_t)b * M;
    for (int t = 0; t < M; ++t) e[t] = energy[(size_t)b * M + t];
    const int slen = min((int)src_lens[b], S);
    int pos = 0;
    for (int i = 0; i < slen; ++i) {
        const int d = (int)dur[(size_t)b * S + i];
        float v = 0.f;
        if (d > 0) {
            float acc = 0.f;
            for (int t = pos; t < pos + d && t < M; ++t) acc += e[t];
            v = acc / (float)d;
        }
        if (i < M) e[i] = v;
        pos += d;
    }
    for (int i = 0; i < S; ++i) out[(size_t)b * S + i] = (i < slen && i < M) ? e[i] : 0.f;
}


// ---------------------------------------------------------------------------------------------
// Bidirectional single-layer GRU recurrence (nn.GRU, gates r|z|n; modules.py:620-640), one CTA per (utterance, direction).
//   gi  = x W_ih^T + b_ih          precomputed for every step by the dense GEMM          [B, T, 3H]
//   gh  = h W_hh^T + b_hh;  r = s(gi_r + gh_r); z = s(gi_z + gh_z); n = tanh(gi_n + r * gh_n);  h' = (1-z) n + z h
// The recurrent weights live transposed in shared memory ([H][3H], 192 KiB at H = 128) for the whole sequence.
// Like the reference, the recurrence runs over ALL T steps (padded phonemes included: no packing).
__global__ void gru_bidir_kernel(const float* __restrict__ gi_f, const float* __restrict__ gi_b,
                                 const float* __restrict__ whh_f, const float* __restrict__ bhh_f,
                                 const float* __restrict__ whh_b, const float* __restrict__ bhh_b, int T, int H,
                                 float* __restrict__ out, float* __restrict__ h_final) {
    CTTS_PDL_SYNC();
    extern __shared__ float sm[];
    const int G = 3 * H;
    float* wt = sm;              // [H][G]
    float* hs = wt + (size_t)H * G;   // [H]
    float* gh = hs + H;          // [G]
    const int b = blockIdx.x, dir = blockIdx.y;
    const float* whh = dir ? whh_b : whh_f;
    const float* bhh = dir ? bhh_b : bhh_f;
    const float* gi = (dir ? gi_b : gi_f) + (size_t)b * T * G;
    const int j = threadIdx.x;   // 0 .. G-1
    for (int i = j; i < H * G; i += G) {
        const int row = i / H, k = i - row * H;      // whh[row][k]
        wt[(size_t)k * G + row] = whh[i];
    }
    if (j < H) hs[j] = 0.f;
    const float bj = bhh[j];
    __syncthreads();
    for (int step = 0; step < T; ++step) {
        const int t = dir ? (T - 1 - step) : step;
        float acc = bj;
#pragma unroll 8
        for (int k = 0; k < H; ++k) acc = fmaf(wt[(size_t)k * G + j], hs[k], acc);
        gh[j] = acc;
        __syncthreads();
        float hn = 0.f;
        if (j < H) {
            const float* g = gi + (size_t)t * G;
            const float r = 1.f / (1.f + expf(-(g[j] + gh[j])));
            const float z = 1.f / (1.f + expf(-(g[H + j] + gh[H + j])));
            const float n = tanhf(g[2 * H + j] + r * gh[2 * H + j]);
            hn = (1.f - z) * n + z * hs[j];
            out[((size_t)b * T + t) * (2 * H) + dir * H + j] = hn;
        }
        __syncthreads();
        if (j < H) hs[j] = hn;
        __syncthreads();
    }
    if (j < H) h_final[(size_t)b * 2 * H + dir * H + j] = hs[j];
}

// y[r, n] = sum_{k<K} x[r,k] w[n,k] + bias[n] (+ residual[r,n]); tiny K (the 4-d phoneme prosody code, modules.py:861)
__global__ void linear_smallk_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                     const float* __restrict__ residual, size_t rows, int K, int N, float* __restrict__ y) {
    CTTS_PDL_SYNC();
    const size_t total = rows * (size_t)N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / N;
        const int n = (int)(i - r * N);
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(x[r * K + k], w[(size_t)n * K + k], acc);
        if (bias) acc += bias[n];
        if (residual) acc += residual[i];
        y[i] = acc;
    }
}

}  // namespace ctts

// =============================================================================================
// C ABI
// =============================================================================================
using namespace ctts;

template <int DH>
static int launch_attention(const float* qkv, const int64_t* lens, int B, int T, int C, int H, float scale, float* out,
                            cudaStream_t st) {
    const size_t sm = (size_t)(3 * 32 * (DH + 4) + 32 * 33) * sizeof(float);
    ensure_smem(attention_fp32_kernel<DH>, sm);
    dim3 grid((T + 31) / 32, H, B);
    launch_k(attention_fp32_kernel<DH>, grid, 128, sm, st, qkv, lens, T, C, scale, out);
    return check_launch("attention");
}

extern "C" {

int ctts_abi_version(void) { return CTTS_ABI_VERSION; }
const char* ctts_last_error(void) { return g_err; }

int ctts_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("cudaGetDevice failed (no CUDA device?)"); return -1; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}

int ctts_embed_tokens(const int64_t* tokens, const float* table, const float* pe, int pe_rows, float embed_scale,
                      int B, int S, int C, int vocab, float* x, float* word, const int64_t* lens, int pos_mode,
                      void* stream) {
    CTTS_REQUIRE(B > 0 && S > 0 && C % 4 == 0, "embed_tokens: bad shape B=%d S=%d C=%d", B, S, C);
    CTTS_REQUIRE(pe_rows > S - (pos_mode ? 1 : 0), "embed_tokens: positional table has %d rows, need > %d", pe_rows, S);
    CTTS_REQUIRE(S <= 32 * POS_MAXCH, "embed_tokens: S=%d too long (max %d)", S, 32 * POS_MAXCH);
    dim3 grid(B, (S + EMB_ROWS - 1) / EMB_ROWS);
    launch_k(embed_tokens_kernel, grid, 256, 0, (cudaStream_t)stream, tokens, table, pe, embed_scale, S, C, vocab, x, word, lens,
                                                              pos_mode);
    return check_launch("embed_tokens");
}

int ctts_add_positions(const float* x, const float* pe, int pe_rows, const float* alpha, const int64_t* lens, int B, int T,
                       int C, int pos_mode, float* y, void* stream) {
    CTTS_REQUIRE(y != nullptr && y != x, "add_positions: y must be a separate buffer (CTAs re-read x[..., 0] of earlier rows)");
    CTTS_REQUIRE(B > 0 && T > 0 && C % 4 == 0, "add_positions: bad shape B=%d T=%d C=%d", B, T, C);
    CTTS_REQUIRE(pe_rows > T - (pos_mode ? 1 : 0), "add_positions: positional table has %d rows, need > %d", pe_rows, T);
    CTTS_REQUIRE(T <= 32 * POS_MAXCH, "add_positions: T=%d too long (max %d)", T, 32 * POS_MAXCH);
    dim3 grid(B, (T + POS_ROWS - 1) / POS_ROWS);
    launch_k(add_positions_kernel, grid, 256, 0, (cudaStream_t)stream, x, pe, alpha, lens, T, C, pos_mode, y);
    return check_launch("add_positions");
}

static int layernorm_impl(const float* x, const float* gamma, const float* beta, float eps, const int64_t* lens, int B,
                          int T, int C, float* y, int np, void* const* planes, void* stream) {
    CTTS_REQUIRE(C % 4 == 0 && C <= 1024, "layernorm: C=%d unsupported (need C %% 4 == 0, C <= 1024)", C);
    const int rows = B * T;
    CTTS_REQUIRE(rows > 0, "layernorm: empty input");
    CTTS_REQUIRE(np >= 0 && np <= 3 && np != 1, "layernorm: n_planes must be 0, 2 or 3");
    const int grid = (rows + 7) / 8;
    PlanePtrs pp{{nullptr, nullptr, nullptr}};
    for (int p = 0; p < np; ++p) {
        CTTS_REQUIRE(planes && planes[p], "layernorm: NULL output plane %d", p);
        pp.p[p] = (__nv_bfloat16*)planes[p];
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (np == 3) launch_k(layernorm_kernel<3>, grid, 256, 0, st, x, gamma, beta, eps, lens, rows, T, C, y, pp);
    else if (np == 2) launch_k(layernorm_kernel<2>, grid, 256, 0, st, x, gamma, beta, eps, lens, rows, T, C, y, pp);
    else launch_k(layernorm_kernel<0>, grid, 256, 0, st, x, gamma, beta, eps, lens, rows, T, C, y, pp);
    return check_launch("layernorm");
}

int ctts_layernorm(const float* x, const float* gamma, const float* beta, float eps, const int64_t* lens, int B, int T,
                   int C, float* y, void* stream) {
    CTTS_REQUIRE(y != nullptr, "layernorm: y is NULL");
    return layernorm_impl(x, gamma, beta, eps, lens, B, T, C, y, 0, nullptr, stream);
}

int ctts_layernorm_planes(const float* x, const float* gamma, const float* beta, float eps, const int64_t* lens, int B, int T,
                          int C, float* y, int n_planes, void* const* planes, void* stream) {
    return layernorm_impl(x, gamma, beta, eps, lens, B, T, C, y, n_planes, planes, stream);
}

int ctts_layernorm_split(const float* x, const float* gamma, const float* beta, float eps, const int64_t* lens, int B,
                         int T, int C, float* y, void* y_hi, void* y_lo, void* stream) {
    CTTS_REQUIRE(y_hi && y_lo, "layernorm_split: y_hi / y_lo are NULL");
    void* planes[3] = {y_hi, y_lo, nullptr};
    return layernorm_impl(x, gamma, beta, eps, lens, B, T, C, y, 2, planes, stream);
}

int ctts_conv1d_gemm(const float* x, const float* w, const float* bias, float alpha, const float* col_scale,
                     const float* col_shift, int act, const float* residual, const int64_t* lens, int B, int T,
                     int Cin, int N, int taps, float* y, void* stream) {
    CTTS_REQUIRE(B > 0 && T > 0 && N > 0 && taps >= 1 && (taps & 1), "conv1d_gemm: bad shape B=%d T=%d N=%d taps=%d", B,
                 T, N, taps);
    CTTS_REQUIRE(Cin % 16 == 0, "conv1d_gemm: Cin=%d must be a multiple of 16", Cin);
    CTTS_REQUIRE((col_scale == nullptr) == (col_shift == nullptr), "conv1d_gemm: col_scale/col_shift must come together");
    if (taps == 1 && (N <= 16 || (long long)B * T <= 32)) {
        const int chunks = (N + SK_NCH - 1) / SK_NCH;
        const long long warps = (long long)B * T * chunks;
        launch_k(skinny_linear_kernel, (unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream, 
            x, w, bias, alpha, col_scale, col_shift, act, residual, lens, (long long)B * T, T, Cin, N, chunks, y);
        return check_launch("conv1d_gemm (skinny)");
    }
    dim3 grid((T + GM - 1) / GM, (N + GN - 1) / GN, B);
    const GAddr ga{1, (long long)T * Cin, 0, Cin, 0, 0, taps * Cin, (long long)T * N, 0, N, 1};
    launch_k(conv1d_gemm_fp32_kernel, grid, 256, 0, (cudaStream_t)stream, x, w, bias, alpha, col_scale, col_shift, act,
                                                                    residual, lens, T, Cin, N, taps, y, ga);
    return check_launch("conv1d_gemm");
}

int ctts_batched_gemm_fp32(const float* x, const float* w, float alpha, const int64_t* lens, int lens_div, int Z, int mod,
                           int T, int K, int N, long long x_so, long long x_sh, int x_ld, long long w_so, long long w_sh,
                           int w_ld, long long y_so, long long y_sh, int y_ld, float* y, void* stream) {
    CTTS_REQUIRE(Z > 0 && mod > 0 && T > 0 && N > 0 && K > 0 && K % 16 == 0, "batched_gemm_fp32: bad shape Z=%d T=%d K=%d N=%d",
                 Z, T, K, N);
    CTTS_REQUIRE(x_ld % 4 == 0 && w_ld % 4 == 0 && x_so % 4 == 0 && x_sh % 4 == 0 && w_so % 4 == 0 && w_sh % 4 == 0,
                 "batched_gemm_fp32: operand strides must be multiples of 4 floats");
    dim3 grid((T + GM - 1) / GM, (N + GN - 1) / GN, Z);
    const GAddr ga{mod, x_so, x_sh, x_ld, w_so, w_sh, w_ld, y_so, y_sh, y_ld, lens_div > 0 ? lens_div : 1};
    launch_k(conv1d_gemm_fp32_kernel, grid, 256, 0, (cudaStream_t)stream, x, w, nullptr, alpha, nullptr, nullptr, CTTS_ACT_NONE,
                                                                    nullptr, lens, T, K, N, 1, y, ga);
    return check_launch("batched_gemm_fp32");
}

int ctts_pack_conv_weight(const float* w, int N, int Cin, int taps, float* packed, void* stream) {
    const size_t total = (size_t)N * Cin * taps;
    CTTS_REQUIRE(total > 0, "pack_conv_weight: empty");
    const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    launch_k(pack_conv_weight_kernel, grid, 256, 0, (cudaStream_t)stream, w, N, Cin, taps, packed);
    return check_launch("pack_conv_weight");
}

int ctts_attention(const float* qkv, const int64_t* lens, int B, int T, int C, int H, float scale, float* out,
                   void* stream) {
    CTTS_REQUIRE(B > 0 && T > 0 && H > 0 && C % H == 0, "attention: bad shape B=%d T=%d C=%d H=%d", B, T, C, H);
    CTTS_REQUIRE(lens != nullptr, "attention: lens is NULL");
    const int dh = C / H;
    switch (dh) {
        case 128: return launch_attention<128>(qkv, lens, B, T, C, H, scale, out, (cudaStream_t)stream);
        case 64: return launch_attention<64>(qkv, lens, B, T, C, H, scale, out, (cudaStream_t)stream);
        case 32: return launch_attention<32>(qkv, lens, B, T, C, H, scale, out, (cudaStream_t)stream);
        default: set_error("attention: head_dim %d unsupported (32, 64, 128)", dh); return 2;
    }
}

int ctts_decode_durations(const float* log_d, float d_control, int n, float* dur, void* stream) {
    CTTS_REQUIRE(n > 0, "decode_durations: empty");
    launch_k(decode_durations_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, log_d, d_control, n, dur);
    return check_launch("decode_durations");
}

int ctts_length_scan(const float* dur_f32, const int64_t* dur_i64, const int64_t* src_lens, int B, int S,
                     int32_t* cum_lr, int32_t* cum_m2p, int64_t* mel_len, void* stream) {
    CTTS_REQUIRE((dur_f32 != nullptr) != (dur_i64 != nullptr), "length_scan: exactly one of dur_f32 / dur_i64");
    CTTS_REQUIRE(B > 0 && S > 0, "length_scan: bad shape");
    // mel_len has room for 2*B entries: [0,B) = LR lengths, [B,2B) = mel2ph lengths
    launch_k(length_scan_kernel, B, 32, 0, (cudaStream_t)stream, dur_f32, dur_i64, src_lens, S, cum_lr, cum_m2p, mel_len,
                                                           mel_len + B);
    return check_launch("length_scan");
}

int ctts_length_expand(const float* src, const float* table, const int64_t* row_index, const int32_t* cum_lr, int B,
                       int S, int C, int M, int accumulate, float* out, const int32_t* cum_m2p, int64_t* mel2ph, int M2,
                       void* stream) {
    CTTS_REQUIRE((src != nullptr) != (table != nullptr), "length_expand: exactly one of src / table");
    CTTS_REQUIRE(table == nullptr || row_index != nullptr, "length_expand: table needs row_index");
    CTTS_REQUIRE(C % 4 == 0 && M > 0, "length_expand: bad shape C=%d M=%d", C, M);
    const int rows = M > M2 ? M : (mel2ph ? M2 : M);
    dim3 grid((rows + 7) / 8, B);
    launch_k(length_expand_kernel, grid, 256, 0, (cudaStream_t)stream, src, table, row_index, cum_lr, S, C, M, accumulate, out,
                                                                 cum_m2p, mel2ph, M2);
    return check_launch("length_expand");
}

int ctts_cwt_to_pitch(const float* cwt, int cwt_stride, const float* scale_w, const float* mean, const float* std,
                      int stat_stride, float std_scale, float eps, const float* uv_src, int use_uv, int B, int T,
                      float* f0_norm, float* f0_denorm, int64_t* pitch_idx, void* stream) {
    CTTS_REQUIRE(B > 0 && T > 1, "cwt_to_pitch: need T > 1 (unbiased std), got B=%d T=%d", B, T);
    CTTS_REQUIRE(cwt_stride >= 10 && (uv_src || !use_uv || cwt_stride >= 11), "cwt_to_pitch: cwt_stride=%d", cwt_stride);
    const size_t sm = (size_t)T * sizeof(float);
    CTTS_REQUIRE(sm <= 200 * 1024, "cwt_to_pitch: T=%d too long", T);
    ensure_smem(cwt_to_pitch_kernel, sm);
    launch_k(cwt_to_pitch_kernel, B, 256, sm, (cudaStream_t)stream, cwt, cwt_stride, scale_w, mean, std, stat_stride, std_scale,
                                                              eps, uv_src, use_uv, T, f0_norm, f0_denorm, pitch_idx);
    return check_launch("cwt_to_pitch");
}

int ctts_f0_to_pitch(const float* f0_norm, const float* uv_src, int n, float* f0_denorm, int64_t* pitch_idx,
                     void* stream) {
    CTTS_REQUIRE(n > 0, "f0_to_pitch: empty");
    launch_k(f0_to_pitch_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, f0_norm, uv_src, n, f0_denorm, pitch_idx);
    return check_launch("f0_to_pitch");
}

int ctts_frame_pitch(float* pred, int ldp, const float* f0_target, const float* uv_target, const int64_t* mel2ph, int use_uv,
                     int n, float* f0_out, float* f0_denorm, int64_t* pitch_idx, void* stream) {
    CTTS_REQUIRE((pred || f0_target) && mel2ph && f0_out && f0_denorm && pitch_idx && n > 0, "frame_pitch: bad arguments");
    CTTS_REQUIRE(!use_uv || uv_target || (pred && ldp >= 2), "frame_pitch: use_uv needs a uv target or a 2-wide prediction");
    launch_k(frame_pitch_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, pred, ldp, f0_target, uv_target, mel2ph, use_uv, n,
             f0_out, f0_denorm, pitch_idx);
    return check_launch("frame_pitch");
}

int ctts_gather_index(const int64_t* idx_ph, const int64_t* mel2ph, int B, int S, int M, int64_t* out, void* stream) {
    CTTS_REQUIRE(idx_ph && mel2ph && out && B > 0 && S > 0 && M > 0, "gather_index: bad arguments");
    const int n = B * M;
    launch_k(gather_index_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, idx_ph, mel2ph, S, M, n, out);
    return check_launch("gather_index");
}

int ctts_phoneme_pitch(const float* f0, const int64_t* mel2ph, const int64_t* src_lens, const int64_t* mel_lens, int B, int S,
                       int M, float* out, void* stream) {
    CTTS_REQUIRE(f0 && mel2ph && src_lens && mel_lens && out && B > 0 && S > 0 && M > 0, "phoneme_pitch: bad arguments");
    dim3 grid((S + 7) / 8, B);
    launch_k(phoneme_pitch_kernel, grid, 256, 0, (cudaStream_t)stream, f0, mel2ph, src_lens, mel_lens, S, M, out);
    return check_launch("phoneme_pitch");
}

int ctts_gather_add(const float* table, const int64_t* idx, int rows, int C, int table_rows, float* x, void* stream) {
    CTTS_REQUIRE(rows > 0 && C % 4 == 0, "gather_add: bad shape");
    launch_k(gather_add_kernel, (rows + 7) / 8, 256, 0, (cudaStream_t)stream, table, idx, rows, C, table_rows, x);
    return check_launch("gather_add");
}

int ctts_bucketize(const float* v, float v_scale, const float* bins, int n_bins, int n, int64_t* idx, void* stream) {
    CTTS_REQUIRE(n > 0 && n_bins > 0, "bucketize: bad shape");
    launch_k(bucketize_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, v, v_scale, bins, n_bins, n, idx);
    return check_launch("bucketize");
}

int ctts_add_row_broadcast(const float* x, const float* row, int B, int T, int C, float* y, void* stream) {
    CTTS_REQUIRE(C % 4 == 0, "add_row_broadcast: C %% 4 != 0");
    const size_t total4 = (size_t)B * T * (C / 4);
    const int grid = (int)((total4 + 255) / 256 < 8192 ? (total4 + 255) / 256 : 8192);
    launch_k(add_row_broadcast_kernel, grid, 256, 0, (cudaStream_t)stream, x, row, T, C, total4, y);
    return check_launch("add_row_broadcast");
}

int ctts_split_planes(const float* x, size_t n, int n_planes, void* const* planes, void* stream) {
    CTTS_REQUIRE(n > 0 && (n_planes == 2 || n_planes == 3) && planes, "split_planes: bad arguments");
    PlanePtrs pp{{nullptr, nullptr, nullptr}};
    uintptr_t align = reinterpret_cast<uintptr_t>(x);
    for (int p = 0; p < n_planes; ++p) {
        CTTS_REQUIRE(planes[p], "split_planes: NULL plane %d", p);
        pp.p[p] = (__nv_bfloat16*)planes[p];
        align |= reinterpret_cast<uintptr_t>(planes[p]);
    }
    const int vec = (align & 15) == 0 ? 1 : 0;
    const size_t work = vec ? (n + 7) / 8 : n;
    const int grid = (int)((work + 255) / 256 < 8192 ? (work + 255) / 256 : 8192);
    if (n_planes == 3) launch_k(split_bf16_kernel<3>, grid, 256, 0, (cudaStream_t)stream, x, n, pp, vec);
    else launch_k(split_bf16_kernel<2>, grid, 256, 0, (cudaStream_t)stream, x, n, pp, vec);
    return check_launch("split_planes");
}

int ctts_split_bf16(const float* x, size_t n, void* hi, void* lo, void* stream) {
    void* planes[3] = {hi, lo, nullptr};
    return ctts_split_planes(x, n, 2, planes, stream);
}

int ctts_fastformer_pool(const float* logits, const float* values, const int64_t* lens, int B, int T, int heads, int head_size,
                         float* pooled, void* stream) {
    CTTS_REQUIRE(B > 0 && T > 0 && heads > 0 && head_size >= 1 && head_size <= 4, "fastformer_pool: bad shape heads=%d hs=%d",
                 heads, head_size);
    CTTS_REQUIRE(lens != nullptr, "fastformer_pool: lens is NULL");
    dim3 grid(B, (heads + 31) / 32);
    // the reference divides by python's attention_head_size ** 0.5 (fastformer.py:310,328)
    launch_k(fastformer_pool_kernel, grid, 256, 0, (cudaStream_t)stream, logits, values, lens, T, heads, head_size,
                                                                   (float)sqrt((double)head_size), pooled);
    return check_launch("fastformer_pool");
}

int ctts_binary(const float* a, const float* b, int op, int b_rowwise, const int64_t* lens, int B, int T, int C, float* y,
                void* stream) {
    CTTS_REQUIRE(C % 4 == 0 && B > 0 && T > 0 && (op == 0 || op == 1), "binary: bad arguments");
    const size_t total4 = (size_t)B * T * (C / 4);
    const int grid = (int)((total4 + 255) / 256 < 8192 ? (total4 + 255) / 256 : 8192);
    launch_k(binary_kernel, grid, 256, 0, (cudaStream_t)stream, a, b, op, b_rowwise, lens, T, C, total4, y);
    return check_launch("binary");
}

int ctts_glu(const float* h, int rows, int C, float* g, void* stream) {
    CTTS_REQUIRE(rows > 0 && C > 0, "glu: bad shape");
    const size_t total = (size_t)rows * C;
    const int grid = (int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
    launch_k(glu_kernel, grid, 256, 0, (cudaStream_t)stream, h, C, (size_t)rows, g);
    return check_launch("glu");
}

int ctts_dwconv_bn_swish(const float* g, const float* w, int K, const float* scale, const float* shift, int B, int T, int C,
                         float* y, void* stream) {
    CTTS_REQUIRE(B > 0 && T > 0 && C > 0 && (K & 1), "dwconv_bn_swish: bad shape");
    dim3 grid((C + 127) / 128, T, B);
    launch_k(dwconv_bn_swish_kernel, grid, 128, 0, (cudaStream_t)stream, g, w, K, scale, shift, T, C, y);
    return check_launch("dwconv_bn_swish");
}

int ctts_relshift_softmax(const float* content, const float* pos, int Z, int T, int ldp, float sqrt_dim, float* P,
                          void* stream) {
    CTTS_REQUIRE(Z > 0 && T > 0 && ldp >= T, "relshift_softmax: bad shape");
    const size_t rows = (size_t)Z * T;
    launch_k(relshift_softmax_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, content, pos, T, ldp, sqrt_dim, rows,
                                                                                        P);
    return check_launch("relshift_softmax");
}

int ctts_relshift_softmax_planes(const float* content, const float* pos, int Z, int T, int ld, int ldp, float sqrt_dim,
                                 int n_planes, void* const* planes, void* stream) {
    CTTS_REQUIRE(content && pos && planes && Z > 0 && T > 0 && ld >= T && ldp >= T && (n_planes == 2 || n_planes == 3),
                 "relshift_softmax_planes: bad arguments");
    PlanePtrs pp{{nullptr, nullptr, nullptr}};
    for (int p = 0; p < n_planes; ++p) {
        CTTS_REQUIRE(planes[p], "relshift_softmax_planes: NULL plane");
        pp.p[p] = (__nv_bfloat16*)planes[p];
    }
    const size_t rows = (size_t)Z * T;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (n_planes == 3) launch_k(relshift_softmax_planes_kernel<3>, grid, 256, 0, (cudaStream_t)stream, content, pos, T, ld, ldp, sqrt_dim, rows, pp);
    else launch_k(relshift_softmax_planes_kernel<2>, grid, 256, 0, (cudaStream_t)stream, content, pos, T, ld, ldp, sqrt_dim, rows, pp);
    return check_launch("relshift_softmax_planes");
}

int ctts_pad_heads_planes(const float* x, const float* bias, int rows, int ld_in, int c0, int H, int DH, int DHp, int n_planes,
                          void* const* planes, void* stream) {
    CTTS_REQUIRE(x && planes && rows > 0 && H > 0 && DH > 0 && DHp >= DH && (n_planes == 2 || n_planes == 3),
                 "pad_heads_planes: bad arguments");
    PlanePtrs pp{{nullptr, nullptr, nullptr}};
    for (int p = 0; p < n_planes; ++p) {
        CTTS_REQUIRE(planes[p], "pad_heads_planes: NULL plane");
        pp.p[p] = (__nv_bfloat16*)planes[p];
    }
    const size_t total = (size_t)rows * H * DHp;
    const int grid = (int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
    if (n_planes == 3) launch_k(pad_heads_planes_kernel<3>, grid, 256, 0, (cudaStream_t)stream, x, bias, ld_in, c0, H, DH, DHp, total, pp);
    else launch_k(pad_heads_planes_kernel<2>, grid, 256, 0, (cudaStream_t)stream, x, bias, ld_in, c0, H, DH, DHp, total, pp);
    return check_launch("pad_heads_planes");
}

int ctts_transpose_heads(const float* x, int B, int T, int ld_in, int c0, int H, int DH, int ldt, float* xt, void* stream) {
    CTTS_REQUIRE(B > 0 && T > 0 && H > 0 && DH > 0 && ldt >= T, "transpose_heads: bad shape");
    dim3 grid((ldt + 31) / 32, (DH + 31) / 32, B * H);
    launch_k(transpose_heads_kernel, grid, 256, 0, (cudaStream_t)stream, x, T, ld_in, c0, H, DH, ldt, xt);
    return check_launch("transpose_heads");
}

int ctts_aligner_attention(const float* q, const float* k, const float* prior, const int64_t* src_lens, float temperature, int B,
                           int M, int S, int C, float* soft, float* logprob, void* stream) {
    CTTS_REQUIRE(B > 0 && M > 0 && S > 0 && C > 0 && src_lens && prior, "aligner_attention: bad arguments");
    const size_t sm = ((size_t)C * (S + 1) + 8 * (size_t)C) * sizeof(float);
    CTTS_REQUIRE(sm <= 200 * 1024, "aligner_attention: S=%d too long for the shared-memory key tile", S);
    ensure_smem(aligner_attention_kernel, sm);
    dim3 grid((M + 7) / 8, B);
    launch_k(aligner_attention_kernel, grid, 256, sm, (cudaStream_t)stream, q, k, prior, src_lens, temperature, M, S, C, soft,
                                                                      logprob);
    return check_launch("aligner_attention");
}

int ctts_mas(const float* attn, const int64_t* src_lens, const int64_t* mel_lens, int B, int M, int S, uint8_t* prev_workspace,
             float* hard, float* dur, void* stream) {
    CTTS_REQUIRE(B > 0 && M > 0 && S > 0 && src_lens && mel_lens && prev_workspace, "mas: bad arguments");
    const size_t sm = 2 * (size_t)S * sizeof(float);
    CTTS_REQUIRE(sm <= 48 * 1024, "mas: S=%d too long", S);
    const int threads = S >= 1024 ? 1024 : ((S + 31) / 32) * 32;
    launch_k(mas_kernel, B, threads, sm, (cudaStream_t)stream, attn, src_lens, mel_lens, M, S, prev_workspace, hard, dur);
    return check_launch("mas");
}

int ctts_phoneme_energy(const float* dur, const int64_t* src_lens, const float* energy, int B, int S, int M, float* workspace,
                        float* out, void* stream) {
    CTTS_REQUIRE(B > 0 && S > 0 && M > 0 && workspace, "phoneme_energy: bad arguments");
    launch_k(phoneme_energy_kernel, (B + 31) / 32, 32, 0, (cudaStream_t)stream, dur, src_lens, energy, B, S, M, workspace, out);
    return check_launch("phoneme_energy");
}

int ctts_gru_bidir(const float* gi_fwd, const float* gi_bwd, const float* w_hh_fwd, const float* b_hh_fwd,
                   const float* w_hh_bwd, const float* b_hh_bwd, int B, int T, int H, float* out, float* h_final,
                   void* stream) {
    CTTS_REQUIRE(B > 0 && T > 0 && H > 0 && 3 * H <= 1024, "gru_bidir: bad shape B=%d T=%d H=%d", B, T, H);
    const size_t sm = ((size_t)H * 3 * H + H + 3 * H) * sizeof(float);
    CTTS_REQUIRE(sm <= 227 * 1024, "gru_bidir: hidden size %d does not fit the shared-memory weight tile", H);
    ensure_smem(gru_bidir_kernel, sm);
    launch_k(gru_bidir_kernel, dim3(B, 2), 3 * H, sm, (cudaStream_t)stream, gi_fwd, gi_bwd, w_hh_fwd, b_hh_fwd, w_hh_bwd, b_hh_bwd,
                                                                      T, H, out, h_final);
    return check_launch("gru_bidir");
}

int ctts_linear_smallk(const float* x, const float* w, const float* bias, const float* residual, int rows, int K, int N,
                       float* y, void* stream) {
    CTTS_REQUIRE(rows > 0 && K > 0 && K <= 64 && N > 0, "linear_smallk: bad shape");
    const size_t total = (size_t)rows * N;
    const int grid = (int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
    launch_k(linear_smallk_kernel, grid, 256, 0, (cudaStream_t)stream, x, w, bias, residual, (size_t)rows, K, N, y);
    return check_launch("linear_smallk");
}

}  // extern "C"
