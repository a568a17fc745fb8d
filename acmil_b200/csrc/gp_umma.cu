// tcgen05 / TMA / TMEM implementation of the fused gated-attention pool pass (sm_100a).
// Placeholder until the kernel lands: reports "unsupported" so ACMIL_IMPL_AUTO picks the FFMA path.
#include "gp_common.cuh"

int gp_umma_supported(const acmil_gp_shape&) { return 0; }
int gp_umma_pack(const acmil_gp_shape&, const acmil_gp_weights&, unsigned char*, cudaStream_t) { return ACMIL_OK; }
int gp_launch_main_umma(const GpMainParams&, cudaStream_t) {
  acmil_set_error("tcgen05 kernel not built");
  return ACMIL_E_UNSUPPORTED;
}
