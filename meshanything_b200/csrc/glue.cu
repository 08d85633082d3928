// glue.cu -- the small data-movement / elementwise kernels around the GEMM and attention kernels of the
// encoder (a1-a8) and the detokenizer (a17-a18).  All HBM-bound, 8- or 16-byte vector accesses.
#include "canon.cuh"
#include "internal.h"
#include "internal_enc.h"

namespace ma {

// a1: FourierEmbedder.forward (embedder.py:87-105) + cat with the normals (sal_perceiver.py:87-89), rounded
// to fp16 as the input_proj Linear does under autocast.  Row layout [xyz(3) | sin (coord-major, 8 freqs) 24 |
// cos 24 | normals 3 | zero padding to 256] so that the canonical Linear (K % 256 == 0) can consume it.
__global__ void fourier_embed_kernel(const __half* __restrict__ pc, long rows, __half* __restrict__ out) {
  const long row = (long)blockIdx.x * (blockDim.x / 64) + threadIdx.x / 64;
  const int c = threadIdx.x & 63;  // 64 threads per row, each writes 4 columns
  if (row >= rows) return;
  const __half* p = pc + row * 6;
  __half o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int col = 4 * c + k;
    float v = 0.0f;
    if (col < 3) {
      v = __half2float(p[col]);
    } else if (col < 27) {
      const int j = col - 3;
      v = sinf(__half2float(p[j / 8]) * (float)(1 << (j % 8)));
    } else if (col < 51) {
      const int j = col - 27;
      v = cosf(__half2float(p[j / 8]) * (float)(1 << (j % 8)));
    } else if (col < 54) {
      v = __half2float(p[3 + col - 51]);
    }
    o[k] = __float2half_rn(v);
  }
  *reinterpret_cast<uint2*>(out + row * 256 + 4 * c) = *reinterpret_cast<uint2*>(o);
}

int launch_fourier_embed(const __half* pc, long rows, __half* out, cudaStream_t st) {
  fourier_embed_kernel<<<(unsigned)((rows + 3) / 4), 256, 0, st>>>(pc, rows, out);
  count_launch();
  return check_launch("fourier_embed_kernel") ? 0 : 1;
}

// Copies head slices of a [rows][ld] fp16 matrix into a head-major buffer:
//   dst[((slot*H + h)*T + t)*64 + d] = src[m*ld + col0 + h*head_stride + d],  slot = m / rows_per_slot, t = m % rows_per_slot
// (rows_per_slot = 1, T = 1 gives the row-major [rows][H*64] layout the attention kernel reads q from).
__global__ void scatter_heads_kernel(const __half* __restrict__ src, int ld, int col0, int head_stride, int H,
                                     int rows_per_slot, long T, __half* __restrict__ dst, long rows) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte piece each
  const long per_row = (long)H * 8;
  if (idx >= rows * per_row) return;
  const long m = idx / per_row;
  const int r = (int)(idx % per_row), h = r / 8, piece = r % 8;
  const long slot = m / rows_per_slot, t = m % rows_per_slot;
  const uint4 u = *reinterpret_cast<const uint4*>(src + m * ld + col0 + h * head_stride + 8 * piece);
  *reinterpret_cast<uint4*>(dst + ((slot * H + h) * T + t) * 64 + 8 * piece) = u;
}

int launch_scatter_heads(const __half* src, int ld, int col0, int head_stride, int H, int rows_per_slot, long T,
                         __half* dst, long rows, cudaStream_t st) {
  const long n = rows * H * 8;
  scatter_heads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, ld, col0, head_stride, H, rows_per_slot, T, dst,
                                                                    rows);
  count_launch();
  return check_launch("scatter_heads_kernel") ? 0 : 1;
}

// x32 += float(y16)   (fp32 residual stream: x + attn / x + mlp, transformer_blocks.py:110-111,224-225)
__global__ void residual_add_f32_kernel(float* __restrict__ x, const __half* __restrict__ y, long n4) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<float4*>(x)[i];
  const uint2 u = reinterpret_cast<const uint2*>(y)[i];
  const __half2* h = reinterpret_cast<const __half2*>(&u);
  const float2 p = __half22float2(h[0]), q = __half22float2(h[1]);
  a.x += p.x; a.y += p.y; a.z += q.x; a.w += q.y;
  reinterpret_cast<float4*>(x)[i] = a;
}
// x16 = fp16(float(x16) + float(y16))  (fp16 residual stream of the 16 "transformer" blocks: its input,
// the post_kl output, is fp16 under autocast -- SURVEY.md 8a a4)
__global__ void residual_add_f16_kernel(__half* __restrict__ x, const __half* __restrict__ y, long n4) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  uint2 ux = reinterpret_cast<uint2*>(x)[i];
  const uint2 uy = reinterpret_cast<const uint2*>(y)[i];
  __half2* hx = reinterpret_cast<__half2*>(&ux);
  const __half2* hy = reinterpret_cast<const __half2*>(&uy);
  hx[0] = __hadd2(hx[0], hy[0]);
  hx[1] = __hadd2(hx[1], hy[1]);
  reinterpret_cast<uint2*>(x)[i] = ux;
}
int launch_residual_add(float* x32, __half* x16, const __half* y, long n, cudaStream_t st) {
  const long n4 = n / 4;
  if (x32) residual_add_f32_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x32, y, n4);
  else residual_add_f16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x16, y, n4);
  count_launch();
  return check_launch("residual_add_kernel") ? 0 : 1;
}

// generic strided row copy with conversion; cols % 4 == 0
template <typename S, typename D>
__global__ void convert_rows_kernel(const S* __restrict__ src, long lds, D* __restrict__ dst, long ldd, long rows,
                                    int cols, long src_rows_mod) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = cols / 4;
  if (idx >= rows * c4) return;
  const long r = idx / c4;
  const int c = (int)(idx % c4) * 4;
  const long rs = src_rows_mod > 0 ? r % src_rows_mod : r;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float v = (float)src[rs * lds + c + k];
    dst[r * ldd + c + k] = (D)v;
  }
}
int launch_convert_rows(const void* src, int src_f16, long lds, void* dst, int dst_f16, long ldd, long rows, int cols,
                        long src_rows_mod, cudaStream_t st) {
  const long n = rows * (cols / 4);
  const unsigned g = (unsigned)((n + 255) / 256);
  if (src_f16 && dst_f16)
    convert_rows_kernel<__half, __half><<<g, 256, 0, st>>>((const __half*)src, lds, (__half*)dst, ldd, rows, cols, src_rows_mod);
  else if (src_f16)
    convert_rows_kernel<__half, float><<<g, 256, 0, st>>>((const __half*)src, lds, (float*)dst, ldd, rows, cols, src_rows_mod);
  else if (dst_f16)
    convert_rows_kernel<float, __half><<<g, 256, 0, st>>>((const float*)src, lds, (__half*)dst, ldd, rows, cols, src_rows_mod);
  else
    convert_rows_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, rows, cols, src_rows_mod);
  count_launch();
  return check_launch("convert_rows_kernel") ? 0 : 1;
}

// out32[r] = (mask && !mask[r] ? 0 : float(y16[r])) + table[r % table_rows]    (width 768)
__global__ void add_table_kernel(const __half* __restrict__ y16, const int* __restrict__ mask,
                                 const float* __restrict__ table, int table_rows, float* __restrict__ out, long rows) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 each, 192 per row
  if (idx >= rows * 192) return;
  const long r = idx / 192;
  const int c = (int)(idx % 192) * 4;
  const uint2 u = *reinterpret_cast<const uint2*>(y16 + r * 768 + c);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
  float2 p = __half22float2(h[0]), q = __half22float2(h[1]);
  if (mask && !mask[r]) p = q = make_float2(0.f, 0.f);
  const float4 t = *reinterpret_cast<const float4*>(table + (long)(r % table_rows) * 768 + c);
  *reinterpret_cast<float4*>(out + r * 768 + c) = make_float4(p.x + t.x, p.y + t.y, q.x + t.z, q.y + t.w);
}
int launch_add_table(const __half* y16, const int* mask, const float* table, int table_rows, float* out, long rows,
                     cudaStream_t st) {
  const long n = rows * 192;
  add_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y16, mask, table, table_rows, out, rows);
  count_launch();
  return check_launch("add_table_kernel") ? 0 : 1;
}

// a9 post-processing + a17 get_codes (meshanything.py:142,163-172,178-212):
// gen_ids [B][max_new] raw generate() output (position 0 = the predicted bos, dropped; the last position is
// dropped too); token u of face f, vertex v, quantizer q = gen_ids[1 + 9f + 3v + q];  specials {0,1,2} -> absent,
// others id-3.  code16[b][f][v*1024 + d] = fp16((c0 + c1) + c2) with absent codes = 0 ; mask[b][f] = all 9 present.
__global__ void gather_codes_kernel(const int32_t* __restrict__ gen_ids, int max_new, int F,
                                    const float* __restrict__ codebook, __half* __restrict__ code16,
                                    int* __restrict__ mask, int32_t* __restrict__ ids_out) {
  const long bf = blockIdx.x;  // b*F + f
  const long b = bf / F;
  const int f = (int)(bf % F);
  __shared__ int tok[9];
  if (threadIdx.x < 9) {
    const int pos = 1 + 9 * f + threadIdx.x;
    int t = (pos < max_new - 1) ? gen_ids[b * max_new + pos] : 1;  // beyond the kept range = eos (never happens for 9F+2)
    t = (t < 3) ? -1 : t - 3;
    tok[threadIdx.x] = t;
    if (ids_out) ids_out[bf * 9 + threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int all = 1;
    for (int i = 0; i < 9; i++) all &= (tok[i] >= 0);
    mask[bf] = all;
  }
  // 3 vertices x 1024 dims = 3072 outputs, 256 threads x 12
  for (int o = threadIdx.x; o < 3072; o += blockDim.x) {
    const int v = o / 1024, d = o % 1024;
    float c[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const int t = tok[3 * v + q];
      c[q] = t >= 0 ? codebook[(long)t * 1024 + d] : 0.0f;
    }
    code16[bf * 3072 + o] = __float2half_rn((c[0] + c[1]) + c[2]);
  }
}
int launch_gather_codes(const int32_t* gen_ids, int max_new, int B, int F, const float* codebook, __half* code16,
                        int* mask, int32_t* ids_out, cudaStream_t st) {
  gather_codes_kernel<<<(unsigned)((long)B * F), 256, 0, st>>>(gen_ids, max_new, F, codebook, code16, mask, ids_out);
  count_launch();
  return check_launch("gather_codes_kernel") ? 0 : 1;
}

// to_coor_logits argmax + undiscretize (meshanything.py:69-78,214-223): logits16 [B*F][9*128] ->
// xyz [B*F][9] = bin/128 - 0.5 ; faces with mask 0 -> NaN.  One warp per (face, coordinate).
__global__ void coords_kernel(const __half* __restrict__ logits, const int* __restrict__ mask, float* __restrict__ xyz,
                              long faces) {
  const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= faces * 9) return;
  const long f = w / 9;
  const __half* p = logits + w * 128;
  float bv = -INFINITY;
  int bi = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int i = lane + 32 * k;
    const float v = __half2float(p[i]);
    if (v > bv) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) xyz[w] = mask[f] ? ((float)bi / 128.0f - 0.5f) : __int_as_float(0x7fc00000);
}
int launch_coords(const __half* logits, const int* mask, float* xyz, long faces, cudaStream_t st) {
  const long threads = faces * 9 * 32;
  coords_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(logits, mask, xyz, faces);
  count_launch();
  return check_launch("coords_kernel") ? 0 : 1;
}

}  // namespace ma
