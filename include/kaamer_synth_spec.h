/*
 * kaamer_synth_spec.h — the counter-based synthetic workload of the C4 configuration
 * (BASELINE.json configs[3]: UniRef90-bacteria-scale DB, ~50 M proteins / ~15 G residues), as pure
 * integer functions of (seed, record index, position).  SURVEY.md §7 "Scale of config 4": the
 * database is generated on the GPU, never written as FASTA, and the CPU oracle checks a sampled query
 * subset by streaming THE SAME generator.  This header is that generator: it is compiled into the
 * CUDA generator (kaamer_b200/csrc/synth.cu) and into its CPU twin (oracle/synth_oracle.cpp); a third,
 * independent numpy/Python statement in tests/test_synth_spec.py pins both.
 *
 * This is workload-generation code (bench / test infrastructure), not part of the reference's path.
 *
 * Model (same as kaamer_b200/synth.py but counter-based):
 *   record i:  meta = Philox(i, 0, STREAM_META) = (m0, m1, m2, m3)
 *     member(i)    = i > 0 && m0 < 2^30                      (25 % of the records are family members)
 *     founder(i)   = member ? nearest non-member j' <= j = (m2 * i) >> 32 : i   (a founder that precedes i)
 *     length(i)    = LEN_Q[m1(founder(i)) >> 22]             (lognormal median 290, sigma 0.6, [30, 5000])
 *     sub_thr(i)   = 5 % .. 30 % as a u32 threshold from m3
 *     residue(i,p) = member && Philox(i, p/4, STREAM_SUB)[p%4] < sub_thr(i)
 *                      ? letter(Philox(i, p/4, STREAM_SUBRES)[p%4])
 *                      : letter(Philox(founder(i), p/4, STREAM_RES)[p%4])
 *   query j of batch b: record t = (q0 * n_proteins) >> 32 with 10 % i.i.d. substitutions
 *   letter(u) = Swiss-Prot composition over the 20 standard letters (AA_THR).
 */
#ifndef KAAMER_SYNTH_SPEC_H
#define KAAMER_SYNTH_SPEC_H
#include <stdint.h>

#include "kaamer_synth_tables.h"

#if defined(__CUDACC__)
#define KSYN_FN __host__ __device__ __forceinline__
#else
#define KSYN_FN static inline
#endif

enum {
  KSYN_STREAM_META = 1,
  KSYN_STREAM_RES = 2,
  KSYN_STREAM_SUB = 3,
  KSYN_STREAM_SUBRES = 4,
  KSYN_STREAM_QMETA = 5,
  KSYN_STREAM_QSUB = 6,
  KSYN_STREAM_QSUBRES = 7
};
#define KSYN_MEMBER_THR 0x40000000u /* 25 % */
#define KSYN_QUERY_SUB_THR 0x1999999Au /* 10 % */

typedef struct ksyn_u4 {
  uint32_t x, y, z, w;
} ksyn_u4;

KSYN_FN void ksyn_mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo) {
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}

/* Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1) */
KSYN_FN ksyn_u4 ksyn_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    ksyn_mulhilo(0xD2511F53u, c0, &h0, &l0);
    ksyn_mulhilo(0xCD9E8D57u, c2, &h1, &l1);
    const uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  ksyn_u4 o;
  o.x = c0;
  o.y = c1;
  o.z = c2;
  o.w = c3;
  return o;
}

/* counter layout: c0 = record (low 32), c1 = block (position / 4), c2 = record (high) | batch << 16, c3 = stream */
KSYN_FN ksyn_u4 ksyn_draw(uint64_t seed, uint64_t rec, uint32_t block, uint32_t batch, uint32_t stream) {
  return ksyn_philox((uint32_t)rec, block, (uint32_t)(rec >> 32) | (batch << 16), stream, (uint32_t)seed,
                     (uint32_t)(seed >> 32));
}
KSYN_FN uint32_t ksyn_lane(ksyn_u4 v, uint32_t p) {
  const uint32_t k = p & 3u;
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

/* letter index 0..19 into KAAMER_SYNTH_LETTERS: first k with u < thr[k] */
KSYN_FN uint32_t ksyn_letter_index(const uint32_t *thr, uint32_t u) {
  uint32_t k = 0;
  for (int i = 0; i < 19; ++i) k += u >= thr[i] ? 1u : 0u;
  return k;
}

typedef struct ksyn_meta {
  uint64_t founder;  /* == record index for a founder */
  uint32_t length;
  uint32_t sub_thr;  /* 0 for a founder */
} ksyn_meta;

KSYN_FN int ksyn_is_member(uint64_t seed, uint64_t i) {
  return i > 0 && ksyn_draw(seed, i, 0, 0, KSYN_STREAM_META).x < KSYN_MEMBER_THR;
}

KSYN_FN ksyn_meta ksyn_record_meta(uint64_t seed, uint64_t i, const uint16_t *len_q) {
  ksyn_meta m;
  const ksyn_u4 d = ksyn_draw(seed, i, 0, 0, KSYN_STREAM_META);
  if (i > 0 && d.x < KSYN_MEMBER_THR) {
    uint64_t j = ((uint64_t)d.z * i) >> 32; /* 0 .. i-1 */
    while (ksyn_is_member(seed, j)) --j;    /* record 0 is a founder */
    m.founder = j;
    m.length = len_q[ksyn_draw(seed, j, 0, 0, KSYN_STREAM_META).y >> 22];
    /* 5 % + u * 25 %: 0x0CCCCCCD = 0.05 * 2^32, 0x40000000 = 0.25 * 2^32 */
    m.sub_thr = 0x0CCCCCCDu + (uint32_t)(((uint64_t)d.w * 0x40000000ull) >> 32);
  } else {
    m.founder = i;
    m.length = len_q[d.y >> 22];
    m.sub_thr = 0;
  }
  return m;
}

/* residues [p0, p0+4) of record i (p0 % 4 == 0) as letter indices packed in 4 bytes (little-endian) */
KSYN_FN uint32_t ksyn_record_block(uint64_t seed, uint64_t i, const ksyn_meta *m, uint32_t block,
                                   const uint32_t *thr) {
  const ksyn_u4 f = ksyn_draw(seed, m->founder, block, 0, KSYN_STREAM_RES);
  uint32_t l0 = ksyn_letter_index(thr, f.x), l1 = ksyn_letter_index(thr, f.y), l2 = ksyn_letter_index(thr, f.z),
           l3 = ksyn_letter_index(thr, f.w);
  if (m->sub_thr) {
    const ksyn_u4 s = ksyn_draw(seed, i, block, 0, KSYN_STREAM_SUB);
    if (s.x < m->sub_thr || s.y < m->sub_thr || s.z < m->sub_thr || s.w < m->sub_thr) {
      const ksyn_u4 r = ksyn_draw(seed, i, block, 0, KSYN_STREAM_SUBRES);
      if (s.x < m->sub_thr) l0 = ksyn_letter_index(thr, r.x);
      if (s.y < m->sub_thr) l1 = ksyn_letter_index(thr, r.y);
      if (s.z < m->sub_thr) l2 = ksyn_letter_index(thr, r.z);
      if (s.w < m->sub_thr) l3 = ksyn_letter_index(thr, r.w);
    }
  }
  return l0 | (l1 << 8) | (l2 << 16) | (l3 << 24);
}

/* query j of batch b: the record it is sampled from */
KSYN_FN uint64_t ksyn_query_record(uint64_t seed, uint64_t n_proteins, uint64_t j, uint32_t batch) {
  const ksyn_u4 d = ksyn_draw(seed, j, 0, batch, KSYN_STREAM_QMETA);
  /* 64-bit fraction (d.x, d.y) of n_proteins (n_proteins < 2^32) */
  const uint64_t u = ((uint64_t)d.x << 32) | d.y;
  const uint64_t hi = (u >> 32) * n_proteins, lo = ((u & 0xFFFFFFFFull) * n_proteins) >> 32;
  return (hi + lo) >> 32;
}
/* residues [4*block, 4*block+4) of query j: the record's, with 10 % i.i.d. substitutions */
KSYN_FN uint32_t ksyn_query_block(uint64_t seed, uint64_t j, uint32_t batch, uint64_t rec, const ksyn_meta *m,
                                  uint32_t block, const uint32_t *thr) {
  uint32_t w = ksyn_record_block(seed, rec, m, block, thr);
  const ksyn_u4 s = ksyn_draw(seed, j, block, batch, KSYN_STREAM_QSUB);
  if (s.x < KSYN_QUERY_SUB_THR || s.y < KSYN_QUERY_SUB_THR || s.z < KSYN_QUERY_SUB_THR || s.w < KSYN_QUERY_SUB_THR) {
    const ksyn_u4 r = ksyn_draw(seed, j, block, batch, KSYN_STREAM_QSUBRES);
    if (s.x < KSYN_QUERY_SUB_THR) w = (w & 0xFFFFFF00u) | ksyn_letter_index(thr, r.x);
    if (s.y < KSYN_QUERY_SUB_THR) w = (w & 0xFFFF00FFu) | (ksyn_letter_index(thr, r.y) << 8);
    if (s.z < KSYN_QUERY_SUB_THR) w = (w & 0xFF00FFFFu) | (ksyn_letter_index(thr, r.z) << 16);
    if (s.w < KSYN_QUERY_SUB_THR) w = (w & 0x00FFFFFFu) | (ksyn_letter_index(thr, r.w) << 24);
  }
  return w;
}

#endif
