// placeholder until the tcgen05 engine lands (next commit): reports "unsupported" so AUTO picks SIMT.
#include "epilogue.cuh"
namespace anyloc {
bool gemm_tc_supported(const float*, const float*, int, const float*, const float*, int, int, int, int,
                       const EpiParams&) { return false; }
int gemm_tc_launch(const float*, const float*, int, const float*, const float*, int, int, int, int,
                   const EpiParams&, cudaStream_t) {
  set_error("tcgen05 GEMM engine not built");
  return ANYLOC_ERR_UNSUPPORTED;
}
}  // namespace anyloc
