// synth.cu — device-side generator of the synthetic C4 workload (include/kaamer_synth_spec.h).
//
// Workload generation for bench.py / tests (SURVEY.md §7 "Scale of config 4": ~15 G residues are
// generated on the GPU with a counter-based RNG and never pass through the host); not part of the
// reference's path.  The parity checks stream a CPU twin of this generator (test infrastructure); both
// include the same integer specification.
#include "internal.cuh"

#include "../../include/kaamer_synth_spec.h"

namespace kaamer {

__constant__ uint32_t c_aa_thr[20] = KAAMER_SYNTH_AA_THR;
__constant__ uint16_t c_len_q[1024] = KAAMER_SYNTH_LEN_Q;
__constant__ char c_letters[21] = KAAMER_SYNTH_LETTERS;

__global__ void k_synth_record_lengths(uint64_t seed, uint64_t first, uint64_t n, uint32_t *len) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  len[i] = ksyn_record_meta(seed, first + i, c_len_q).length;
}

__global__ void k_synth_query_lengths(uint64_t seed, uint64_t n_proteins, uint32_t batch, uint64_t first, uint64_t n,
                                      uint32_t *len) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t rec = ksyn_query_record(seed, n_proteins, first + i, batch);
  len[i] = ksyn_record_meta(seed, rec, c_len_q).length;
}

// one warp per record / query: lanes stride over the 4-residue blocks
template <bool QUERY>
__global__ void __launch_bounds__(256) k_synth_residues(uint64_t seed, uint64_t n_proteins, uint32_t batch,
                                                        uint64_t first, uint64_t n,
                                                        const uint64_t *__restrict__ seq_off,
                                                        uint8_t *__restrict__ res) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (w >= n) return;
  uint64_t rec = first + w;
  if (QUERY) rec = ksyn_query_record(seed, n_proteins, first + w, batch);
  const ksyn_meta m = ksyn_record_meta(seed, rec, c_len_q);
  uint8_t *dst = res + seq_off[w];
  const uint32_t nblocks = (m.length + 3) / 4;
  for (uint32_t b = lane; b < nblocks; b += 32) {
    const uint32_t v = QUERY ? ksyn_query_block(seed, first + w, batch, rec, &m, b, c_aa_thr)
                             : ksyn_record_block(seed, rec, &m, b, c_aa_thr);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (b * 4 + k < m.length) dst[b * 4 + k] = (uint8_t)c_letters[(v >> (8 * k)) & 0xFF];
  }
}

}  // namespace kaamer

using namespace kaamer;

extern "C" {

int kaamer_synth_record_lengths(const kaamer_synth_cfg *cfg, uint64_t first, uint64_t n, uint32_t *d_len,
                                void *stream) {
  if (!cfg || (n && !d_len)) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  if (n == 0) return KAAMER_OK;
  k_synth_record_lengths<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cfg->seed, first, n, d_len);
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

int kaamer_synth_record_residues(const kaamer_synth_cfg *cfg, uint64_t first, uint64_t n, const uint64_t *d_seq_off,
                                 uint8_t *d_res, void *stream) {
  if (!cfg || (n && (!d_seq_off || !d_res))) {
    set_error("null argument");
    return KAAMER_ERR_ARG;
  }
  if (n == 0) return KAAMER_OK;
  k_synth_residues<false><<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      cfg->seed, cfg->n_proteins, 0, first, n, d_seq_off, d_res);
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

int kaamer_synth_query_lengths(const kaamer_synth_cfg *cfg, uint32_t batch, uint64_t first, uint64_t n,
                               uint32_t *d_len, void *stream) {
  if (!cfg || (n && !d_len) || cfg->n_proteins == 0 || cfg->n_proteins >= (1ull << 32) || batch >= 65536) {
    set_error("bad argument");
    return KAAMER_ERR_ARG;
  }
  if (n == 0) return KAAMER_OK;
  k_synth_query_lengths<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cfg->seed, cfg->n_proteins,
                                                                                         batch, first, n, d_len);
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

int kaamer_synth_query_residues(const kaamer_synth_cfg *cfg, uint32_t batch, uint64_t first, uint64_t n,
                                const uint64_t *d_seq_off, uint8_t *d_res, void *stream) {
  if (!cfg || (n && (!d_seq_off || !d_res)) || cfg->n_proteins == 0 || cfg->n_proteins >= (1ull << 32) ||
      batch >= 65536) {
    set_error("bad argument");
    return KAAMER_ERR_ARG;
  }
  if (n == 0) return KAAMER_OK;
  k_synth_residues<true><<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      cfg->seed, cfg->n_proteins, batch, first, n, d_seq_off, d_res);
  KCUDA(cudaGetLastError());
  return KAAMER_OK;
}

}  // extern "C"
