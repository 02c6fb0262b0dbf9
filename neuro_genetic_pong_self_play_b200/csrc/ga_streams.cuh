// ga_streams.cuh -- Philox4x32-10 counter layout of the GA kernels (K4), shared by ngp_ops.cu and ngp_population.cu.
// key = (seed lo, seed hi); counter = (c0, c1, generation, stream tag):
//   STREAM_INIT   c0,c1 = low/high word of the 4-gene block index q        words 0..3 -> genes 4q..4q+3           (ga.py:85-87)
//   STREAM_SELECT c0 = offspring slot, c1 = block b                         word w -> draw 4b+w, index = hi32(word * n)  (ga.py:94)
//   STREAM_CXDO   c0 = pair, c1 = 0                                         word 0 -> u < cxpb                     (varAnd)
//   STREAM_CXU    c0 = pair, c1 = gene / 4                                  word gene % 4 -> u of cxBlend          (ga.py:89)
//   STREAM_MUTDO  c0 = individual, c1 = 0                                   word 0 -> u < mutpb                    (varAnd)
//   STREAM_MUTU   c0 = individual, c1 = gene / 4                            word gene % 4 -> u < indpb             (ga.py:91-92)
//   STREAM_MUTZ   c0 = individual, c1 = gene                                words 0,1 -> Box-Muller normal
// uniforms: u = (word >> 8) * 2^-24 in [0,1).  Results are independent of launch geometry.
#pragma once
#include <stdint.h>
#include "policy.cuh"

enum : uint32_t { STREAM_SELECT = 0x53454C31u, STREAM_CXDO = 0x43584431u, STREAM_CXU = 0x43585531u, STREAM_MUTDO = 0x4D544431u,
                  STREAM_MUTU = 0x4D545531u, STREAM_MUTZ = 0x4D545A31u, STREAM_INIT = 0x494E4931u };

__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }   // [0,1)

// one gene of cxBlend (DEAP tools.cxBlend; every FP32 operation individually rounded)
__device__ __forceinline__ void blend_gene(float alpha, float u, float &x0, float &x1)
{
    const float gamma = __fsub_rn(__fmul_rn((float)(1.0 + 2.0 * (double)alpha), u), alpha);
    const float one_m = __fsub_rn(1.0f, gamma);
    const float c0 = __fadd_rn(__fmul_rn(one_m, x0), __fmul_rn(gamma, x1));
    const float c1 = __fadd_rn(__fmul_rn(gamma, x0), __fmul_rn(one_m, x1));
    x0 = c0; x1 = c1;
}

// standard normal of gene `gene` of individual `ind`: Box-Muller on two words of a per-gene Philox block
__device__ __forceinline__ float mut_normal(uint32_t ind, uint32_t gene, uint32_t gen, uint32_t k0, uint32_t k1)
{
    uint32_t o[4];
    pol::philox4x32(ind, gene, gen, STREAM_MUTZ, k0, k1, o);
    const float u1 = ((float)(o[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);      // (0,1]
    const float u2 = u01(o[1]);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
