// stubs.cu — entry points whose kernels are not written yet (replaced as rows of
// SURVEY §8 land).  They fail loudly; nothing here computes on the CPU.
#include "internal.cuh"
using namespace kaamer;
extern "C" {
int kaamer_gpu_search_nucleotide(kaamer_gpu_t *, const uint8_t *, const uint64_t *, uint32_t, const kaamer_opts *,
                                 kaamer_hits **) {
  set_error("kaamer_gpu_search_nucleotide: not implemented in this build");
  return KAAMER_ERR_ARG;
}
int kaamer_gpu_get_orfs(kaamer_gpu_t *, const uint8_t *, const uint64_t *, uint32_t, kaamer_orfs **) {
  set_error("kaamer_gpu_get_orfs: not implemented in this build");
  return KAAMER_ERR_ARG;
}
void kaamer_gpu_free_orfs(kaamer_orfs *) {}
int kaamer_gpu_align(kaamer_gpu_t *, const uint8_t *, const uint64_t *, const uint32_t *, const uint32_t *, uint32_t,
                     const kaamer_aln_opts *, kaamer_aln *) {
  set_error("kaamer_gpu_align: not implemented in this build");
  return KAAMER_ERR_ARG;
}
}
