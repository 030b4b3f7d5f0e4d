// TEMPORARY: placeholders until tt_gemm_sm100.cu / tt_scan_sm100.cu land. Every entry fails loudly.
#include "tt_sm100.cuh"
namespace tt {
size_t mlp_sm100_ws_bytes(int, int, int) { return 0; }
int mlp_fwd_sm100(const float*, int, int, int, const float*, const float*, const float*, const float*, float*, float*,
                  int, void*, size_t, cudaStream_t) { set_error("tensor-core projection not built yet"); return 3; }
int mlp_bwd_sm100(const float*, const float*, const float*, const float*, const float*, int, int, int, float*, float*,
                  float*, float*, float*, int, int, void*, size_t, cudaStream_t) { set_error("tensor-core projection not built yet"); return 3; }
size_t step_sm100_ws_bytes(int, int, int, int) { return 0; }
int step_sm100(const StepSm100&, cudaStream_t) { set_error("tensor-core step not built yet"); return 3; }
size_t scan_sm100_ws_bytes(int, long long, int, int) { return 0; }
int scan_topk_sm100(const float*, const float*, const void*, const void*, int, long long, int, int, long long, float*,
                    long long*, void*, size_t, cudaStream_t) { set_error("tensor-core scan not built yet"); return 3; }
}  // namespace tt
