// Shared helpers for libctts_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "ctts_b200.h"

namespace ctts {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

// Raise a kernel's dynamic shared-memory limit only when a launch needs more than was granted before (never while a
// graph capture re-runs a launch of the same size: the first, eager pass has already done it).
void ensure_smem_impl(const void* kernel, size_t bytes);
template <typename K>
inline void ensure_smem(K kernel, size_t bytes) {
    ensure_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// Every kernel of the library is launched with programmaticStreamSerialization: the grid may be scheduled (and run its
// set-up: barrier init, TMEM allocation, descriptor prefetch) while its predecessor in the stream is still draining.
// CTTS_PDL_SYNC() is the first thing a kernel does that touches global memory: it lets the NEXT kernel start launching
// (griddepcontrol.launch_dependents) and then waits until the PREVIOUS grid has completed and its writes are visible
// (griddepcontrol.wait).  Every kernel must execute it, also one that needs nothing from its predecessor: completion
// then chains through the stream exactly as without the attribute.  Opt-in (CTTS_PDL=1): measured with the benchmark
// under CUDA-graph replay it changes nothing (3.59 vs 3.58 ms per step) -- the graph already pre-stages the launches --
// so the default launches without the attribute (the two instructions are no-ops then).
#define CTTS_PDL_SYNC()                                                      \
    do {                                                                     \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      \
        asm volatile("griddepcontrol.wait;" ::: "memory");                   \
    } while (0)

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in check_launch()
}

#define CTTS_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            ctts::set_error(__VA_ARGS__);  \
            return 2;                      \
        }                                  \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact (erf) GELU, tanh etc. -- the reference uses F.gelu (erf form), torch.tanh, relu, x*sigmoid(x)
__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case CTTS_ACT_RELU: return fmaxf(v, 0.f);
        case CTTS_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case CTTS_ACT_TANH: return tanhf(v);
        case CTTS_ACT_SWISH: return v / (1.f + expf(-v));
        default: return v;
    }
}

// ---------------------------------------------------------------------------------------------
// Positions of rows [r0, r1) of one utterance, by the whole CTA: pos[t] = flag(t) ? #flags in [0, t] : 0
// (make_positions, utils/tools.py:640-652, with padding_idx 0).  Every warp ballots 32-row chunks of [0, r1) into
// s_mask; each row then sums the popcounts of the chunks before its own.  r1 <= 32 * POS_MAXCH.
constexpr int POS_MAXCH = 512;
template <class F>
__device__ __forceinline__ void block_positions(int r0, int r1, int* s_pos, unsigned* s_mask, F flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int nch = (r1 + 31) >> 5;
    for (int ch = warp; ch < nch; ch += nw) {
        const int t = ch * 32 + lane;
        const bool f = (t < r1) && flag(t);
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_mask[ch] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < r1 - r0; i += blockDim.x) {
        const int t = r0 + i, ch = t >> 5, l = t & 31;
        const unsigned m = s_mask[ch];
        int p = 0;
        if ((m >> l) & 1u) {
            p = __popc(m & (0xffffffffu >> (31 - l)));
            for (int c = 0; c < ch; ++c) p += __popc(s_mask[c]);
        }
        s_pos[i] = p;
    }
    __syncthreads();
}

}  // namespace ctts
