// stubs.cu — entry points whose kernels are not written yet (replaced as rows of
// SURVEY §8 land).  They fail loudly; nothing here computes on the CPU.
#include "internal.cuh"
using namespace kaamer;
extern "C" {
int kaamer_gpu_align(kaamer_gpu_t *, const uint8_t *, const uint64_t *, const uint32_t *, const uint32_t *, uint32_t,
                     const kaamer_aln_opts *, kaamer_aln *) {
  set_error("kaamer_gpu_align: not implemented in this build");
  return KAAMER_ERR_ARG;
}
}
