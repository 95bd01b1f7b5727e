// tcgen05 engine for the GRU step (placeholder until the TMEM kernel lands in this file).
#include "common.cuh"

namespace ddfa {

bool gru_tc_available() { return false; }

size_t gru_tc_workspace_bytes(int32_t N, int32_t D) {
  (void)N; (void)D;
  return 16;
}

int gru_tc_step_fwd(const float *, const float *, const int32_t *, const float *, const float *, const float *,
                    const float *, const float *, int32_t, int32_t, float *, float *, void *, size_t, cudaStream_t) {
  set_error("tcgen05 engine not built into this library");
  return DDFA_ERR_UNSUPPORTED;
}

}  // namespace ddfa
