// placeholder until the tcgen05 engine lands (next commit)
#include "ctts_common.cuh"
extern "C" int ctts_gemm_bf16x3(const void*, const void*, const void*, const void*, const float*, float, const float*,
                                const float*, int, const float*, const int64_t*, int, int, int, int, int, float*, void*,
                                void*, void*) {
    ctts::set_error("ctts_gemm_bf16x3: tensor-core engine not built in this revision");
    return 3;
}
