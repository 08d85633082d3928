// canon.cuh -- device-side pieces of the canonical arithmetic.
//
// Every floating-point operation whose result is part of the parity contract is spelled with an
// explicit round-to-nearest intrinsic (__fmaf_rn / __fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn /
// __fsqrt_rn) so that nvcc can neither contract a*b+c into an FMA nor split one.  The CPU oracle
// (oracle/decoder_oracle.c, built with -ffp-contract=off) performs the same operations in the same
// order with scalar loops over virtual lanes.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ma_canon_constants.h"

namespace ma {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// exp(x) for x <= 0 (softmax arguments); see include/ma_canon_constants.h
__device__ __forceinline__ float ma_exp(float x) {
  if (x < MA_EXP_FLUSH) return 0.0f;
  float y = fmul(x, MA_LOG2E);
  float n = rintf(y);
  float f = fsub(y, n);
  float p = MA_EXP2_C6;
  p = ffma(p, f, MA_EXP2_C5);
  p = ffma(p, f, MA_EXP2_C4);
  p = ffma(p, f, MA_EXP2_C3);
  p = ffma(p, f, MA_EXP2_C2);
  p = ffma(p, f, MA_EXP2_C1);
  p = ffma(p, f, MA_EXP2_C0);
  int bits = __float_as_int(p) + (((int)n) << 23);
  return __int_as_float(bits);
}

// 8 fp16 values of a 16-byte load -> fp32 (exact)
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// Mixed-precision FMA of sm_100 (PTX fma.rn.f32.f16 -> SASS FHFMA, operands taken straight from the packed halves with
// .H0/.H1 selectors): d = float(a) * float(b) + c with ONE rounding -- the same value as ffma(unpacked a, unpacked b, c),
// because fp16 -> fp32 is exact -- without the 2 conversion instructions per product.  Checked on the B200 against
// convert + FFMA on 67 M random products incl. subnormal, inf and nan inputs: 0 mismatches
// (tools/microbench_cluster.cu, profiles/microbench_cluster_r02.txt), and by the bit-exact GPU suite against the CPU
// oracle.  Default since round 2; -DMA_NO_FHFMA (build.py: MA_B200_NO_FHFMA=1) builds the convert + FFMA variant.
#ifndef MA_NO_FHFMA
#ifndef MA_FHFMA
#define MA_FHFMA 1
#endif
#endif

#ifdef MA_FHFMA
__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(a), "h"(b));
  return c;
}
// acc + sum_j w[j] * x[j], j = 0..7 in this order (the canonical per-lane order), on two 16-byte groups of packed halves
__device__ __forceinline__ float dot8_packed(const uint4& w, const uint4& x, float acc) {
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w}, xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    acc = fhfma((unsigned short)(ww[i] & 0xffffu), (unsigned short)(xw[i] & 0xffffu), acc);
    acc = fhfma((unsigned short)(ww[i] >> 16), (unsigned short)(xw[i] >> 16), acc);
  }
  return acc;
}
#endif

// acc + sum_j w[j] * x[j], j = 0..7 sequentially (the canonical per-lane chain); either build gives the same bits
__device__ __forceinline__ float dot8(const uint4& w, const uint4& x, float acc) {
#ifdef MA_FHFMA
  return dot8_packed(w, x, acc);
#else
  float wf[8], xf[8];
  unpack8(w, wf);
  unpack8(x, xf);
#pragma unroll
  for (int j = 0; j < 8; j++) acc = ffma(wf[j], xf[j], acc);
  return acc;
#endif
}
// o[j] = float(p) * float(v[j]) + o[j], j = 0..7 (the P.V step of the canonical attention: P already rounded to fp16)
__device__ __forceinline__ void pv8(__half p, const uint4& v, float* o) {
#ifdef MA_FHFMA
  const unsigned short ph = __half_as_ushort(p);
  const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    o[2 * i] = fhfma(ph, (unsigned short)(vw[i] & 0xffffu), o[2 * i]);
    o[2 * i + 1] = fhfma(ph, (unsigned short)(vw[i] >> 16), o[2 * i + 1]);
  }
#else
  const float pf = __half2float(p);
  float vf[8];
  unpack8(v, vf);
#pragma unroll
  for (int j = 0; j < 8; j++) o[j] = ffma(pf, vf[j], o[j]);
#endif
}

__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// xor-16,8,4,2,1 butterfly: every lane ends with the canonical sum of the 32 lane partials
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fadd(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Transposing butterfly: each lane holds 32 partials v[0..31] (one per output); on return lane l
// holds in v[0] the canonical (xor-16,8,4,2,1) sum over lanes of partial l.  31 shuffles.
__device__ __forceinline__ float transpose_reduce32(float* v, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; i++) {
      float mine = up ? v[i + s] : v[i];
      float other = up ? v[i] : v[i + s];
      float recv = __shfl_xor_sync(0xffffffffu, other, s);
      v[i] = fadd(mine, recv);
    }
  }
  return v[0];
}

// The same with a chosen lane-bit order.  ORD = 1: xor-4,2,1,8,16 -- the segmented (64-wide) order of the decoder's
// out_proj: the 8 lanes of a 64-element segment are reduced first, then the 4 segments of the 256-wide group as a
// balanced tree.  On return lane l holds accumulator transpose_owner<ORD>(l) in v[0].
template <int ORD>
__device__ __forceinline__ float transpose_reduce32o(float* v, int lane) {
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const int s = 16 >> k;
    const int m = ORD ? (k == 0 ? 4 : k == 1 ? 2 : k == 2 ? 1 : k == 3 ? 8 : 16) : s;
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < s; i++) {
      float mine = up ? v[i + s] : v[i];
      float other = up ? v[i] : v[i + s];
      float recv = __shfl_xor_sync(0xffffffffu, other, m);
      v[i] = fadd(mine, recv);
    }
  }
  return v[0];
}
template <int ORD>
__device__ __forceinline__ int transpose_owner(int lane) {
  if (!ORD) return lane;
  return ((lane & 4) ? 16 : 0) | ((lane & 2) ? 8 : 0) | ((lane & 1) ? 4 : 0) | ((lane & 8) ? 2 : 0) | ((lane & 16) ? 1 : 0);
}
// every lane ends with the xor-4,2,1,8,16 sum (the seg-64 order) of the 32 lane partials
__device__ __forceinline__ float warp_sum_seg64(float v) {
  v = fadd(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fadd(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fadd(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fadd(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = fadd(v, __shfl_xor_sync(0xffffffffu, v, 16));
  return v;
}
// balanced binary tree over 16 values in index order, fed one value at a time (g = 0..15); lv = 4 level registers
__device__ __forceinline__ float tree16_push(float v, int g, float* lv) {
  if (g & 1) {
    v = fadd(lv[0], v);
    if (g & 2) {
      v = fadd(lv[1], v);
      if (g & 4) {
        v = fadd(lv[2], v);
        if (g & 8) v = fadd(lv[3], v); else lv[3] = v;
      } else lv[2] = v;
    } else lv[1] = v;
  } else lv[0] = v;
  return v;   // the total after g = 15
}
// the same over 4 values (g = 0..3): (v0 + v1) + (v2 + v3); lv = 2 registers
__device__ __forceinline__ float tree4_push(float v, int g, float* lv) {
  if (g == 0) lv[0] = v;
  else if (g == 1) lv[0] = fadd(lv[0], v);
  else if (g == 2) lv[1] = v;
  else v = fadd(lv[0], fadd(lv[1], v));
  return v;   // the total after g = 3
}

// pairwise left-to-right tree over n (<= 8) warp sums held in shared memory
__device__ __forceinline__ float warp_tree(const float* s, int n) {
  float b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) b[i] = (i < n) ? s[i] : 0.0f;
  if (n == 8) return fadd(fadd(fadd(b[0], b[1]), fadd(b[2], b[3])), fadd(fadd(b[4], b[5]), fadd(b[6], b[7])));
  if (n == 6) return fadd(fadd(fadd(b[0], b[1]), fadd(b[2], b[3])), fadd(b[4], b[5]));
  if (n == 4) return fadd(fadd(b[0], b[1]), fadd(b[2], b[3]));
  if (n == 2) return fadd(b[0], b[1]);
  // generic (n in {1,3,5,7}): level by level, odd element carried
  int m = n;
  while (m > 1) {
    int k = 0;
    for (int i = 0; i + 1 < m; i += 2) b[k++] = fadd(b[i], b[i + 1]);
    if (m & 1) b[k++] = b[m - 1];
    m = k;
  }
  return b[0];
}

// Canonical block sum over W = 4*blockDim.x values: thread t contributes (x0+x1)+(x2+x3) of its own
// four elements; warp butterfly; warp tree.  `red` = shared scratch of >= 8 floats.  All threads
// return the sum.  Contains two __syncthreads().
__device__ __forceinline__ float block_sum4(float x0, float x1, float x2, float x3, float* red) {
  float p = fadd(fadd(x0, x1), fadd(x2, x3));
  p = warp_sum(p);
  const int nw = blockDim.x >> 5;
  __syncthreads();  // protect `red` against a previous use
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = p;
  __syncthreads();
  return warp_tree(red, nw);
}

// LayerNorm of a row of W = 4*blockDim.x fp32 values, thread t owns elements 4t..4t+3.
__device__ __forceinline__ void layernorm4(float* x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                           float eps, int W, float* red) {
  const float inv = __fdiv_rn(1.0f, (float)W);
  float mean = fmul(block_sum4(x[0], x[1], x[2], x[3], red), inv);
  float d0 = fsub(x[0], mean), d1 = fsub(x[1], mean), d2 = fsub(x[2], mean), d3 = fsub(x[3], mean);
  float var = fmul(block_sum4(fmul(d0, d0), fmul(d1, d1), fmul(d2, d2), fmul(d3, d3), red), inv);
  float rstd = __fdiv_rn(1.0f, __fsqrt_rn(fadd(var, eps)));
  const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * threadIdx.x);
  const float4 b = *reinterpret_cast<const float4*>(beta + 4 * threadIdx.x);
  x[0] = ffma(fmul(d0, rstd), g.x, b.x);
  x[1] = ffma(fmul(d1, rstd), g.y, b.y);
  x[2] = ffma(fmul(d2, rstd), g.z, b.z);
  x[3] = ffma(fmul(d3, rstd), g.w, b.w);
}

// ---- mbarrier / bulk-copy (TMA 1-D) wrappers -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- programmatic dependent launch -----------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace ma
