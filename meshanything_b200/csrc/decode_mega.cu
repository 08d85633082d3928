// decode_mega.cu -- batch-1 greedy decode as ONE persistent kernel: one CTA per SM runs every phase
// of every layer of up to `n_steps` tokens.
//
// Why: a token is 24 x (qkv, attention, out_proj, fc1, fc2) + lm_head = 121 dependent phases that
// together must stream 623.5 MB of weights (+ the KV cache) from HBM in ~100 us.  As separate
// kernels (decode_fast.cu) every phase pays a launch boundary (4.8 us measured even with PDL); with
// grid-wide barriers between phases it still pays barrier + dependent load (2.5 us measured).  Here
//   * each CTA owns a fixed, even-sized block of rows of every weight matrix (contiguous bytes), staged
//     through four shared-memory buffers (qkv 44 KB, out_proj 16 KB, fc1 56 KB, fc2 64 KB) refilled by
//     one bulk async copy (TMA 1-D, mbarrier completion) as soon as the phase that read them ends:
//     the weights of layer L+1 are in flight while layer L computes, so HBM streams continuously;
//   * the K/V rows an SM needs for attention are prefetched into registers before the qkv phase;
//   * there is NO grid barrier: every activation vector is exchanged through L2 as 8-byte words
//     {2 x fp16 (or one fp32), 32-bit epoch}; 8-byte stores are single-copy atomic, so a consumer
//     that sees the epoch sees the data (the NCCL "LL" protocol) -- one L2 round trip per hand-off;
//   * the residual stream lives in shared memory; every CTA recomputes the LayerNorms redundantly.
// Arithmetic is the canonical order of DESIGN.md section 3: results are bit-identical to
// gemm_canon.cu / attention.cu / decode_fast.cu and to the CPU oracle.
#include "canon.cuh"
#include "internal.h"

namespace ma {

constexpr int MG_THREADS = 256;
constexpr int MG_WARPS = 8;
constexpr int PARTF = 66;      // o[64], max, sum
constexpr int MAX_CHUNKS = 72;  // 18432 keys

struct MegaWs {
  uint2 qkv_w[QKV / 2];   // flagged words: {half2, epoch}
  uint2 attn_w[HID / 2];
  uint2 ya_w[HID / 2];    // out_proj output
  uint2 yb_w[HID / 2];    // fc2 output
  uint2 f_w[FFN / 2];
  uint2 cand_w[256 * 2];  // {value bits, epoch}, {index, epoch}
  int error;              // 1: a poll timed out
  int pad_[3];
  unsigned long long trace[1280];
  ma_decoder_weights w;   // device copy of the weight table
  alignas(256) uint2 part_w[NHEAD * MAX_CHUNKS * PARTF];  // {fp32 bits, epoch}
  alignas(256) __half bias_cta[MA_MAX_LAYERS * 160 * 128];  // [layer][cta][128]: this CTA's biases (see BIAS_*)
};

// layout of one CTA's 128 packed biases of a layer
constexpr int BIAS_QKV = 0, BIAS_OUT = 32, BIAS_FC1 = 48, BIAS_FC2 = 112;

struct MegaArgs {
  MegaWs* ws;
  SeqState s;
  __half* kv;  // [layer][kv][head][T][64]   (batch 1)
  long T;
  int n_steps, step_base, max_new, eos_id, pad_id;
  int rows_qkv, rows_out, rows_fc1, rows_fc2, rows_lm;  // rows per CTA of each matrix (even)
  int32_t* out_ids;
  const int32_t* forced;
  __half* logits_out;
  int* all_done;
  int* nkeys_next;
  int trace;
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---- flagged-word exchange -------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint2* p, uint32_t data, uint32_t ep) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(data), "r"(ep) : "memory");
}
__device__ __forceinline__ uint4 ll_load2(const uint2* p) {  // two words
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ll_load1(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
constexpr unsigned SPIN_LIMIT = 1u << 24;  // ~ seconds: never hang the GPU
// spin until both words carry epoch `ep`; returns the two data halves
__device__ __forceinline__ uint2 ll_wait2(const uint2* p, uint32_t ep, int* err) {
  uint4 v = ll_load2(p);
  unsigned spins = 0;
  while (v.y != ep || v.w != ep) {
    if (++spins > SPIN_LIMIT) { *err = 1; break; }
    v = ll_load2(p);
  }
  return make_uint2(v.x, v.z);
}
__device__ __forceinline__ uint32_t ll_wait1(const uint2* p, uint32_t ep, int* err) {
  uint2 v = ll_load1(p);
  unsigned spins = 0;
  while (v.y != ep) {
    if (++spins > SPIN_LIMIT) { *err = 1; break; }
    v = ll_load1(p);
  }
  return v.x;
}
// gather a flagged fp16 vector of `nhalf` elements into shared memory (4 halfs per thread and round)
__device__ __forceinline__ void ll_gather(const uint2* src, int nhalf, uint32_t ep, __half* dst, int* err) {
  for (int u = threadIdx.x; u < nhalf / 4; u += MG_THREADS) {
    const uint2 d = ll_wait2(src + 2 * u, ep, err);
    *reinterpret_cast<uint2*>(dst + 4 * u) = d;
  }
}

// ---- shared memory layout ---------------------------------------------------------------------------
struct alignas(128) MegaSmem {
  uint64_t bar[4];  // full barriers of buffers D (qkv), C (out), A (fc1), B (fc2)
  uint64_t lnbar[2];   // ln1 / ln2 parameter regions
  uint64_t bbar[2];    // packed-bias double buffer
  alignas(16) float ln1[2 * HID];   // gamma | beta of self_attn_layer_norm of the current layer
  alignas(16) float ln2[2 * HID];   // gamma | beta of final_layer_norm
  alignas(16) __half bias[2][128];
  ma_decoder_weights wtab;          // pointer table (kept on chip: every access would be an HBM miss)
  float red[8];
  float wmax[8];
  float ared[8][65];
  float bval[8];
  int bidx[8];
  alignas(16) float hres[HID];  // residual stream
  alignas(16) __half xs[FFN];   // fp16 input vector of the current GEMV
};

__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Rows of this CTA held in shared memory `sw` ([nrows][K], nrows even): pair p = rows (2p, 2p+1); warp w owns
// pairs w, w+8, w+16, w+24 and runs two pairs (4 rows) at a time for instruction-level parallelism.  The four
// warp sums are produced by a partially transposing butterfly (same additions as warp_sum, 6 shuffles for 4
// values).  `emit(n_even, h0, h1)` is called by one lane with the fp16 results of rows n_even, n_even+1.
template <int K, typename Emit>
__device__ __forceinline__ void gemv_pairs(const __half* sw, int nrows, int row0, const __half* bias, const __half* xs,
                                           int warp, int lane, Emit emit) {
  constexpr int G = K / 256;
  const int npairs = nrows >> 1;
#pragma unroll
  for (int i = 0; i < 4; i += 2) {
    const int pA = warp + MG_WARPS * i, pB = pA + MG_WARPS;
    if (pA >= npairs) break;
    const bool hasB = pB < npairs;
    const __half* wA = sw + (size_t)(2 * pA) * K + 8 * lane;
    const __half* wB = sw + (size_t)(2 * (hasB ? pB : pA)) * K + 8 * lane;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
    for (int g = 0; g < G; g++) {
      float xf[8], f[4][8];
      unpack8(*reinterpret_cast<const uint4*>(xs + 256 * g + 8 * lane), xf);
      unpack8(*reinterpret_cast<const uint4*>(wA + 256 * g), f[0]);
      unpack8(*reinterpret_cast<const uint4*>(wA + K + 256 * g), f[1]);
      unpack8(*reinterpret_cast<const uint4*>(wB + 256 * g), f[2]);
      unpack8(*reinterpret_cast<const uint4*>(wB + K + 256 * g), f[3]);
#pragma unroll
      for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int r = 0; r < 4; r++) acc[r] = ffma(f[r][j], xf[j], acc[r]);
      }
    }
    // xor-16: lanes with bit 4 clear keep rows 0,1 (pair A), the others rows 2,3 (pair B)
    const bool up16 = (lane & 16) != 0;
    float k0 = up16 ? acc[2] : acc[0], k1 = up16 ? acc[3] : acc[1];
    float s0 = up16 ? acc[0] : acc[2], s1 = up16 ? acc[1] : acc[3];
    k0 = fadd(k0, __shfl_xor_sync(0xffffffffu, s0, 16));
    k1 = fadd(k1, __shfl_xor_sync(0xffffffffu, s1, 16));
    // xor-8: lanes with bit 3 clear keep the first row of their pair
    const bool up8 = (lane & 8) != 0;
    float k = up8 ? k1 : k0, sx = up8 ? k0 : k1;
    k = fadd(k, __shfl_xor_sync(0xffffffffu, sx, 8));
    k = fadd(k, __shfl_xor_sync(0xffffffffu, k, 4));
    k = fadd(k, __shfl_xor_sync(0xffffffffu, k, 2));
    k = fadd(k, __shfl_xor_sync(0xffffffffu, k, 1));
    // lane 0: row 2pA, lane 8: row 2pA+1, lane 16: row 2pB, lane 24: row 2pB+1
    const int myrow = 2 * (up16 ? pB : pA) + (up8 ? 1 : 0);
    const float bf = bias ? __half2float(bias[min(myrow, nrows - 1)]) : 0.0f;
    const __half hv = __float2half_rn(fadd(k, bf));
    const __half hn = __shfl_down_sync(0xffffffffu, hv, 8);
    if (lane == 0) emit(row0 + 2 * pA, hv, hn);
    if (lane == 16 && hasB) emit(row0 + 2 * pB, hv, hn);
  }
}

__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// refill a weight buffer with rows [row0, row0+nrows) of W[N][K] (one bulk copy, thread 0 only)
__device__ __forceinline__ void refill(__half* dst, const void* W, int row0, int nrows, int K, uint64_t* bar) {
  fence_proxy_async();
  if (nrows > 0) {
    const uint32_t bytes = (uint32_t)nrows * K * 2;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(dst, reinterpret_cast<const __half*>(W) + (size_t)row0 * K, bytes, bar);
  } else {
    mbar_expect_tx(bar, 0);  // plain arrival so that the phase still completes
  }
}

// LayerNorm gamma|beta (2 x 4 KB) and this CTA's 256 bytes of packed biases: small bulk copies issued one layer ahead
__device__ __forceinline__ void fill_ln(float* dst, const float* g, const float* b, uint64_t* bar) {
  fence_proxy_async();
  mbar_expect_tx(bar, 2u * HID * 4);
  bulk_g2s(dst, g, HID * 4, bar);
  bulk_g2s(dst + HID, b, HID * 4, bar);
}
__device__ __forceinline__ void fill_bias(__half* dst, const __half* src, uint64_t* bar) {
  fence_proxy_async();
  mbar_expect_tx(bar, 256);
  bulk_g2s(dst, src, 256, bar);
}

__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(MegaArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  MegaSmem& sm = *reinterpret_cast<MegaSmem*>(smem_raw);
  __half* bufD = reinterpret_cast<__half*>(smem_raw + sizeof(MegaSmem));  // qkv rows
  __half* bufC = bufD + (size_t)a.rows_qkv * HID;                         // out_proj rows
  __half* bufA = bufC + (size_t)a.rows_out * HID;                         // fc1 rows
  __half* bufB = bufA + (size_t)a.rows_fc1 * HID;                         // fc2 rows (K = 4096)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = lane >> 3, li = lane & 7;
  MegaWs* ws = a.ws;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&ws->w);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.wtab);
    for (int i = tid; i < (int)(sizeof(ma_decoder_weights) / 4); i += MG_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const ma_decoder_weights& W = sm.wtab;
  const int NL = W.n_layers;
  const long T = a.T;
  int* err = &ws->error;

  const int cta = blockIdx.x, ncta = gridDim.x;
  const int row0_qkv = cta * a.rows_qkv, n_qkv = max(0, min(a.rows_qkv, QKV - row0_qkv));
  const int row0_out = cta * a.rows_out, n_out = max(0, min(a.rows_out, HID - row0_out));
  const int row0_fc1 = cta * a.rows_fc1, n_fc1 = max(0, min(a.rows_fc1, FFN - row0_fc1));
  const int row0_fc2 = cta * a.rows_fc2, n_fc2 = max(0, min(a.rows_fc2, HID - row0_fc2));
  const int row0_lm = cta * a.rows_lm, n_lm = max(0, min(a.rows_lm, W.vocab - row0_lm));
  // lm rows are staged in D|C|A (contiguous); sub-ranges refilled when each buffer becomes free
  const int lmD = min(n_lm, a.rows_qkv), lmC = max(0, min(n_lm, a.rows_qkv + a.rows_out) - a.rows_qkv),
            lmA = max(0, n_lm - a.rows_qkv - a.rows_out);

  uint32_t parD = 0, parC = 0, parA = 0, parB = 0, parL1 = 0, parL2 = 0, parB0 = 0, parB1 = 0;
  int lc = 0;  // layer instances processed by this launch
  if (tid == 0) {
    for (int i = 0; i < 4; i++) mbar_init(&sm.bar[i], 1);
    for (int i = 0; i < 2; i++) { mbar_init(&sm.lnbar[i], 1); mbar_init(&sm.bbar[i], 1); }
    mbar_fence_init();
    fill_ln(sm.ln1, W.ln1g[0], W.ln1b[0], &sm.lnbar[0]);
    fill_ln(sm.ln2, W.ln2g[0], W.ln2b[0], &sm.lnbar[1]);
    fill_bias(sm.bias[0], ws->bias_cta + ((size_t)0 * 160 + cta) * 128, &sm.bbar[0]);
    fill_bias(sm.bias[1], ws->bias_cta + ((size_t)(NL > 1 ? 1 : 0) * 160 + cta) * 128, &sm.bbar[1]);
    refill(bufD, W.wqkv[0], row0_qkv, n_qkv, HID, &sm.bar[0]);
    refill(bufC, W.wo[0], row0_out, n_out, HID, &sm.bar[1]);
    refill(bufA, W.w1[0], row0_fc1, n_fc1, HID, &sm.bar[2]);
    refill(bufB, W.w2[0], row0_fc2, n_fc2, FFN, &sm.bar[3]);
  }
  __syncthreads();

  // generation state, identical in every CTA
  int pos = a.s.pos[0], gen = a.s.gen[0], tok = a.s.tok[0], fin = a.s.finished[0];
  unsigned long long* tr = (a.trace && cta == 0 && tid == 0) ? ws->trace : nullptr;
  int tri = 0;
#define STAMP() do { if (tr && tri < 1270) tr[tri++] = gtimer(); } while (0)

  for (int step = 0; step < a.n_steps; step++) {
    if (gen >= a.max_new || fin) break;  // uniform across the grid
    const int nkeys = pos + 1;
    const int nch = (nkeys + MA_ATTN_CHUNK - 1) / MA_ATTN_CHUNK;
    const int nitems = nch * NHEAD;
    const uint32_t ep0 = (uint32_t)(a.step_base + step) * (uint32_t)NL + 1u;  // epoch of layer 0 of this step
    STAMP();

    for (int L = 0; L < NL; L++) {
      const uint32_t ep = ep0 + (uint32_t)L;
      const int bsel = lc & 1;            // layer instances alternate between the two bias buffers
      const __half* lb = sm.bias[bsel];   // this layer's packed biases (waited for below)
      __half* kc = a.kv + ((size_t)(L * 2 + 0)) * NHEAD * T * HD;
      __half* vc = a.kv + ((size_t)(L * 2 + 1)) * NHEAD * T * HD;

      // ---------------- K/V prefetch into registers: first attention item of this CTA (rows < pos are old)
      uint4 kreg[8], vreg[8];
      int item = cta;
      if (item < nitems) {
        const int c = item >> 4, h = item & 15;
        const long base = ((long)h * T + (long)c * MA_ATTN_CHUNK) * HD;
#pragma unroll
        for (int rho = 0; rho < 8; rho++) {
          const int r = 32 * rho + 4 * warp + grp;
          if (c * MA_ATTN_CHUNK + r < pos) {
            kreg[rho] = ldcg16(kc + base + (long)r * HD + 8 * li);
            vreg[rho] = ldcg16(vc + base + (long)r * HD + 8 * li);
          }
        }
      }

      // ---------------- qkv phase: input = token embedding (layer 0) or LN2(hres + fc2 output) of the previous layer
      {
        float v[4];
        if (L == 0) {
          float4 X;
          int fidx;
          if (tok < 3) {
            X = *reinterpret_cast<const float4*>(W.extra + (long)tok * HID + 4 * tid);
            fidx = tok;
          } else {
            const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(W.tok_table) +
                                                            (long)(tok - 3) * HID + 4 * tid);
            const __half2* hh = reinterpret_cast<const __half2*>(&u);
            const float2 p0 = __half22float2(hh[0]), p1 = __half22float2(hh[1]);
            X = make_float4(p0.x, p0.y, p1.x, p1.y);
            int r = (gen - 2) % 9;
            if (r < 0) r += 9;
            fidx = r + 3;
          }
          const float4 F = *reinterpret_cast<const float4*>(W.tok_pos + (long)fidx * HID + 4 * tid);
          const float4 C = *reinterpret_cast<const float4*>(W.cond + HID + 4 * tid);
          const float4 P = *reinterpret_cast<const float4*>(W.pos + (long)(pos + 2) * HID + 4 * tid);
          v[0] = fadd(fadd(fadd(X.x, F.x), C.x), P.x);
          v[1] = fadd(fadd(fadd(X.y, F.y), C.y), P.y);
          v[2] = fadd(fadd(fadd(X.z, F.z), C.z), P.z);
          v[3] = fadd(fadd(fadd(X.w, F.w), C.w), P.w);
        } else {
          const float4 hv = *reinterpret_cast<const float4*>(sm.hres + 4 * tid);
          const uint2 d = ll_wait2(ws->yb_w + 2 * tid, ep - 1, err);  // fc2 output of layer L-1
          const __half2* hh = reinterpret_cast<const __half2*>(&d);
          const float2 p0 = __half22float2(hh[0]), p1 = __half22float2(hh[1]);
          v[0] = fadd(hv.x, p0.x); v[1] = fadd(hv.y, p0.y); v[2] = fadd(hv.z, p1.x); v[3] = fadd(hv.w, p1.y);
          mbar_wait(&sm.lnbar[1], parL2);
          parL2 ^= 1;
          layernorm4(v, sm.ln2, sm.ln2 + HID, MA_LN_EPS, HID, sm.red);
          __syncthreads();  // every thread has read its gamma/beta
          if (tid == 0) fill_ln(sm.ln2, W.ln2g[L], W.ln2b[L], &sm.lnbar[1]);
        }
        *reinterpret_cast<float4*>(sm.hres + 4 * tid) = make_float4(v[0], v[1], v[2], v[3]);
        __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(sm.xs + 4 * tid) = u;
      }
      __syncthreads();
      STAMP();
      mbar_wait(&sm.bbar[bsel], bsel ? parB1 : parB0);
      if (bsel) parB1 ^= 1; else parB0 ^= 1;
      mbar_wait(&sm.bar[0], parD);
      parD ^= 1;
      gemv_pairs<HID>(bufD, n_qkv, row0_qkv, lb + BIAS_QKV, sm.xs, warp, lane,
                      [&](int n, __half h0, __half h1) {
                        ll_store(ws->qkv_w + (n >> 1), pack2(h0, h1), ep);
                        if (n >= HID) {  // k / v of the current token also go to the cache for later steps
                          const int e = (n - HID) & (HID - 1), head = e >> 6, d = e & 63;
                          __half* c = (n < 2 * HID) ? kc : vc;
                          *reinterpret_cast<uint32_t*>(c + ((long)head * T + pos) * HD + d) = pack2(h0, h1);
                        }
                      });
      __syncthreads();
      if (tid == 0) {
        if (L + 1 < NL) refill(bufD, W.wqkv[L + 1], row0_qkv, n_qkv, HID, &sm.bar[0]);
        else refill(bufD, W.lm_head, row0_lm, lmD, HID, &sm.bar[0]);
      }
      STAMP();

      // ---------------- attention phase: items (chunk c, head h) = cta, cta + ncta, ...
      for (int it = 0; item < nitems; item += ncta, it++) {
        const int c = item >> 4, h = item & 15;
        const int len = min(MA_ATTN_CHUNK, nkeys - c * MA_ATTN_CHUNK);
        const long base = ((long)h * T + (long)c * MA_ATTN_CHUNK) * HD;
        const int cur = pos - c * MA_ATTN_CHUNK;  // row of the current token inside this chunk (if 0 <= cur < 256)
#pragma unroll
        for (int rho = 0; rho < 8; rho++) {
          const int r = 32 * rho + 4 * warp + grp;
          if (r < len && it > 0 && r != cur) {  // later items were not prefetched
            kreg[rho] = ldcg16(kc + base + (long)r * HD + 8 * li);
            vreg[rho] = ldcg16(vc + base + (long)r * HD + 8 * li);
          }
        }
        // q of this head and, for the chunk that holds it, k / v of the current token: flagged words
        float qf[8];
        {
          const uint2 d0 = ll_wait2(ws->qkv_w + (h * HD + 8 * li) / 2, ep, err);
          const uint2 d1 = ll_wait2(ws->qkv_w + (h * HD + 8 * li) / 2 + 2, ep, err);
          unpack8(make_uint4(d0.x, d0.y, d1.x, d1.y), qf);
        }
        if (cur >= 0 && cur < MA_ATTN_CHUNK) {
          const int rho_c = cur >> 5, gl_c = cur & 31;
          if (4 * warp + grp == gl_c) {
            const uint2 k0 = ll_wait2(ws->qkv_w + (HID + h * HD + 8 * li) / 2, ep, err);
            const uint2 k1 = ll_wait2(ws->qkv_w + (HID + h * HD + 8 * li) / 2 + 2, ep, err);
            const uint2 v0 = ll_wait2(ws->qkv_w + (2 * HID + h * HD + 8 * li) / 2, ep, err);
            const uint2 v1 = ll_wait2(ws->qkv_w + (2 * HID + h * HD + 8 * li) / 2 + 2, ep, err);
#pragma unroll
            for (int rho = 0; rho < 8; rho++)
              if (rho == rho_c) {
                kreg[rho] = make_uint4(k0.x, k0.y, k1.x, k1.y);
                vreg[rho] = make_uint4(v0.x, v0.y, v1.x, v1.y);
              }
          }
        }
        float sreg[8];
        float lmax = -INFINITY;
#pragma unroll
        for (int rho = 0; rho < 8; rho++) {
          const int r = 32 * rho + 4 * warp + grp;
          float kf[8];
          unpack8(kreg[rho], kf);
          float p = 0.0f;
#pragma unroll
          for (int j = 0; j < 8; j++) p = ffma(qf[j], kf[j], p);
          p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 4));
          p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 2));
          p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 1));
          sreg[rho] = fmul(p, 0.125f);
          if (r < len) lmax = fmaxf(lmax, sreg[rho]);
        }
        lmax = warp_max(lmax);
        __syncthreads();  // previous users of wmax / ared are done
        if (lane == 0) sm.wmax[warp] = lmax;
        __syncthreads();
        float cmax = sm.wmax[0];
#pragma unroll
        for (int w2 = 1; w2 < 8; w2++) cmax = fmaxf(cmax, sm.wmax[w2]);
        float l = 0.0f, o[8];
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = 0.0f;
#pragma unroll
        for (int rho = 0; rho < 8; rho++) {
          const int r = 32 * rho + 4 * warp + grp;
          if (r < len) {
            const float e = ma_exp(fsub(sreg[rho], cmax));
            l = fadd(l, e);
            const float pf = __half2float(__float2half_rn(e));
            float vf[8];
            unpack8(vreg[rho], vf);
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = ffma(pf, vf[j], o[j]);
          }
        }
        l = fadd(l, __shfl_xor_sync(0xffffffffu, l, 16));
        l = fadd(l, __shfl_xor_sync(0xffffffffu, l, 8));
#pragma unroll
        for (int j = 0; j < 8; j++) {
          o[j] = fadd(o[j], __shfl_xor_sync(0xffffffffu, o[j], 16));
          o[j] = fadd(o[j], __shfl_xor_sync(0xffffffffu, o[j], 8));
        }
        if (grp == 0) {
#pragma unroll
          for (int j = 0; j < 8; j++) sm.ared[warp][8 * li + j] = o[j];
          if (li == 0) sm.ared[warp][64] = l;
        }
        __syncthreads();
        if (tid < 65) {
          float x[8];
#pragma unroll
          for (int w2 = 0; w2 < 8; w2++) x[w2] = sm.ared[w2][tid];
          const float rsum = fadd(fadd(fadd(x[0], x[1]), fadd(x[2], x[3])), fadd(fadd(x[4], x[5]), fadd(x[6], x[7])));
          uint2* part = ws->part_w + ((long)h * MAX_CHUNKS + c) * PARTF;
          ll_store(part + (tid < 64 ? tid : 65), __float_as_uint(rsum), ep);
          if (tid == 64) ll_store(part + 64, __float_as_uint(cmax), ep);
        }
      }
      // merge of the chunks of head h by the CTA that owns item (chunk 0, head h): ascending order
      if (cta < NHEAD) {
        const int h = cta;
        const uint2* part = ws->part_w + (long)h * MAX_CHUNKS * PARTF;
        __syncthreads();
        // stage {max, sum} of every chunk in shared memory (parallel polls), then each of 64 threads walks its dim
        float* stage = reinterpret_cast<float*>(sm.xs);  // 2 * nch floats
        for (int i = tid; i < 2 * nch; i += MG_THREADS)
          stage[i] = __uint_as_float(ll_wait1(part + (i >> 1) * PARTF + 64 + (i & 1), ep, err));
        __syncthreads();
        if (tid < 64) {
          float M = -INFINITY;
          for (int cc = 0; cc < nch; cc++) M = fmaxf(M, stage[2 * cc]);
          float Lsum = 0.0f, O = 0.0f;
          float oc = __uint_as_float(ll_wait1(part + tid, ep, err));
          for (int cc = 0; cc < nch; cc++) {
            const float onext = (cc + 1 < nch) ? __uint_as_float(ll_wait1(part + (cc + 1) * PARTF + tid, ep, err)) : 0.0f;
            const float wgt = ma_exp(fsub(stage[2 * cc], M));
            Lsum = ffma(stage[2 * cc + 1], wgt, Lsum);
            O = ffma(oc, wgt, O);
            oc = onext;
          }
          const __half r = __float2half_rn(__fdiv_rn(O, Lsum));
          const __half r2 = __shfl_down_sync(0xffffffffu, r, 1);
          if ((tid & 1) == 0) ll_store(ws->attn_w + (h * HD + tid) / 2, pack2(r, r2), ep);
        }
        __syncthreads();
      }
      STAMP();

      // ---------------- out_proj phase
      if (n_out > 0) {
        ll_gather(ws->attn_w, HID, ep, sm.xs, err);
        __syncthreads();
        mbar_wait(&sm.bar[1], parC);
        gemv_pairs<HID>(bufC, n_out, row0_out, lb + BIAS_OUT, sm.xs, warp, lane,
                        [&](int n, __half h0, __half h1) { ll_store(ws->ya_w + (n >> 1), pack2(h0, h1), ep); });
        __syncthreads();
      } else {
        mbar_wait(&sm.bar[1], parC);
      }
      parC ^= 1;
      if (tid == 0) {
        if (L + 1 < NL) refill(bufC, W.wo[L + 1], row0_out, n_out, HID, &sm.bar[1]);
        else refill(bufC, W.lm_head, row0_lm + a.rows_qkv, lmC, HID, &sm.bar[1]);
      }
      STAMP();

      // ---------------- fc1 phase: input = LN1(hres + out_proj)
      {
        const float4 hv = *reinterpret_cast<const float4*>(sm.hres + 4 * tid);
        const uint2 d = ll_wait2(ws->ya_w + 2 * tid, ep, err);
        const __half2* hh = reinterpret_cast<const __half2*>(&d);
        const float2 p0 = __half22float2(hh[0]), p1 = __half22float2(hh[1]);
        float v[4] = {fadd(hv.x, p0.x), fadd(hv.y, p0.y), fadd(hv.z, p1.x), fadd(hv.w, p1.y)};
        mbar_wait(&sm.lnbar[0], parL1);
        parL1 ^= 1;
        layernorm4(v, sm.ln1, sm.ln1 + HID, MA_LN_EPS, HID, sm.red);
        __syncthreads();
        if (tid == 0) fill_ln(sm.ln1, W.ln1g[(L + 1) % NL], W.ln1b[(L + 1) % NL], &sm.lnbar[0]);
        *reinterpret_cast<float4*>(sm.hres + 4 * tid) = make_float4(v[0], v[1], v[2], v[3]);
        __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
        uint2 uo;
        uo.x = *reinterpret_cast<uint32_t*>(&h0);
        uo.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(sm.xs + 4 * tid) = uo;
      }
      __syncthreads();
      mbar_wait(&sm.bar[2], parA);
      parA ^= 1;
      gemv_pairs<HID>(bufA, n_fc1, row0_fc1, lb + BIAS_FC1, sm.xs, warp, lane,
                      [&](int n, __half h0, __half h1) {
                        if (__half2float(h0) < 0.0f) h0 = __float2half_rn(0.0f);
                        if (__half2float(h1) < 0.0f) h1 = __float2half_rn(0.0f);
                        ll_store(ws->f_w + (n >> 1), pack2(h0, h1), ep);
                      });
      __syncthreads();
      if (tid == 0) {
        if (L + 1 < NL) refill(bufA, W.w1[L + 1], row0_fc1, n_fc1, HID, &sm.bar[2]);
        else refill(bufA, W.lm_head, row0_lm + a.rows_qkv + a.rows_out, lmA, HID, &sm.bar[2]);
      }
      STAMP();

      // ---------------- fc2 phase
      if (n_fc2 > 0) {
        ll_gather(ws->f_w, FFN, ep, sm.xs, err);
        __syncthreads();
        mbar_wait(&sm.bar[3], parB);
        gemv_pairs<FFN>(bufB, n_fc2, row0_fc2, lb + BIAS_FC2, sm.xs, warp, lane,
                        [&](int n, __half h0, __half h1) { ll_store(ws->yb_w + (n >> 1), pack2(h0, h1), ep); });
        __syncthreads();
      } else {
        mbar_wait(&sm.bar[3], parB);
      }
      parB ^= 1;
      if (tid == 0) {
        refill(bufB, W.w2[(L + 1 < NL) ? L + 1 : 0], row0_fc2, n_fc2, FFN, &sm.bar[3]);
        // the bias buffer of this layer is free (every warp passed the __syncthreads above or has no fc2 rows):
        // refill it with the biases of the layer that uses it next (L + 2, wrapping into the next token)
      }
      __syncthreads();
      if (tid == 0) fill_bias(sm.bias[bsel], ws->bias_cta + ((size_t)((L + 2) % NL) * 160 + cta) * 128, &sm.bbar[bsel]);
      lc++;
      STAMP();
    }

    // ---------------- lm_head on LN2 of the last layer + greedy pick
    const uint32_t epc = (uint32_t)(a.step_base + step) + 1u;
    {
      const float4 hv = *reinterpret_cast<const float4*>(sm.hres + 4 * tid);
      const uint2 d = ll_wait2(ws->yb_w + 2 * tid, ep0 + (uint32_t)NL - 1u, err);
      const __half2* hh = reinterpret_cast<const __half2*>(&d);
      const float2 p0 = __half22float2(hh[0]), p1 = __half22float2(hh[1]);
      float v[4] = {fadd(hv.x, p0.x), fadd(hv.y, p0.y), fadd(hv.z, p1.x), fadd(hv.w, p1.y)};
      mbar_wait(&sm.lnbar[1], parL2);
      parL2 ^= 1;
      layernorm4(v, sm.ln2, sm.ln2 + HID, MA_LN_EPS, HID, sm.red);
      __syncthreads();
      if (tid == 0) fill_ln(sm.ln2, W.ln2g[0], W.ln2b[0], &sm.lnbar[1]);
      __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
      uint2 uo;
      uo.x = *reinterpret_cast<uint32_t*>(&h0);
      uo.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(sm.xs + 4 * tid) = uo;
    }
    __syncthreads();
    mbar_wait(&sm.bar[0], parD);
    mbar_wait(&sm.bar[1], parC);
    mbar_wait(&sm.bar[2], parA);
    parD ^= 1; parC ^= 1; parA ^= 1;
    float bestv = -INFINITY;
    int besti = 0x7fffffff;
    {
      // n_lm may be odd (vocab 8195): the last pair's second row then reads stale shared memory and is ignored
      const int n_even = (n_lm + 1) & ~1;
      // up to 56 rows = 28 pairs: gemv_pairs covers 4 pairs per warp (32), enough for rows_lm <= 64
      gemv_pairs<HID>(bufD, n_even, row0_lm, nullptr, sm.xs, warp, lane, [&](int n, __half h0, __half h1) {
        const bool two = (n + 1 < row0_lm + n_lm);
        if (a.logits_out) {
          a.logits_out[(long)gen * W.vocab + n] = h0;
          if (two) a.logits_out[(long)gen * W.vocab + n + 1] = h1;
        }
        const float v0 = __half2float(h0);
        if (v0 > bestv || (v0 == bestv && n < besti)) { bestv = v0; besti = n; }
        if (two) {
          const float v1 = __half2float(h1);
          if (v1 > bestv || (v1 == bestv && n + 1 < besti)) { bestv = v1; besti = n + 1; }
        }
      });
    }
    {  // lanes 0 and 16 each tracked the rows they emitted
      const float ov = __shfl_xor_sync(0xffffffffu, bestv, 16);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, 16);
      if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
    }
    if (lane == 0) { sm.bval[warp] = bestv; sm.bidx[warp] = besti; }
    __syncthreads();
    if (tid == 0) {
      float bv = sm.bval[0];
      int bi = sm.bidx[0];
      for (int w2 = 1; w2 < MG_WARPS; w2++)
        if (sm.bval[w2] > bv || (sm.bval[w2] == bv && sm.bidx[w2] < bi)) { bv = sm.bval[w2]; bi = sm.bidx[w2]; }
      __threadfence();  // publish this step's KV-cache rows before the step's final hand-off
      ll_store(ws->cand_w + 2 * cta, __float_as_uint(bv), epc);
      ll_store(ws->cand_w + 2 * cta + 1, (uint32_t)bi, epc);
      // weights of the next token's first layer
      refill(bufD, W.wqkv[0], row0_qkv, n_qkv, HID, &sm.bar[0]);
      refill(bufC, W.wo[0], row0_out, n_out, HID, &sm.bar[1]);
      refill(bufA, W.w1[0], row0_fc1, n_fc1, HID, &sm.bar[2]);
    }
    {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int i = tid; i < ncta; i += MG_THREADS) {
        const uint2 d = ll_wait2(ws->cand_w + 2 * i, epc, err);
        const float v = __uint_as_float(d.x);
        const int ix = (int)d.y;
        if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      __syncthreads();
      if (lane == 0) { sm.bval[warp] = bv; sm.bidx[warp] = bi; }
      __syncthreads();
      bv = sm.bval[0];
      bi = sm.bidx[0];
      for (int w2 = 1; w2 < MG_WARPS; w2++)
        if (sm.bval[w2] > bv || (sm.bval[w2] == bv && sm.bidx[w2] < bi)) { bv = sm.bval[w2]; bi = sm.bidx[w2]; }
      int ntok = bi;
      if (a.forced) ntok = a.forced[gen];
      if (fin) ntok = a.pad_id;
      if (cta == 0 && tid == 0) {
        if (gen < a.max_new) a.out_ids[gen] = ntok;
        if (!fin) a.s.lens[0] = gen + 1;
      }
      if (!fin && ntok == a.eos_id) fin = 1;
      tok = ntok;
      gen += 1;
      pos += 1;
      __syncthreads();
    }
    STAMP();
  }

  // every buffer has a refill in flight here: drain them before the shared memory is released
  mbar_wait(&sm.bar[0], parD);
  mbar_wait(&sm.bar[1], parC);
  mbar_wait(&sm.bar[2], parA);
  mbar_wait(&sm.bar[3], parB);
  mbar_wait(&sm.lnbar[0], parL1);
  mbar_wait(&sm.lnbar[1], parL2);
  mbar_wait(&sm.bbar[0], parB0);
  mbar_wait(&sm.bbar[1], parB1);
  if (cta == 0 && tid == 0) {
    a.s.pos[0] = pos; a.s.gen[0] = gen; a.s.tok[0] = tok; a.s.finished[0] = fin;
    if (a.nkeys_next) *a.nkeys_next = pos + 1;
    if (a.all_done) *a.all_done = fin;
  }
}

// ---- host side -----------------------------------------------------------------------------------
static int g_mega_sms = 0;
size_t mega_workspace_bytes() { return sizeof(MegaWs) + 256; }

static int mega_sms() {
  if (!g_mega_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_mega_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_mega_sms <= 0 || g_mega_sms > 160) g_mega_sms = 148;
  }
  return g_mega_sms;
}
static inline int mega_rpc(int N) { return (((N + mega_sms() - 1) / mega_sms()) + 1) & ~1; }  // even rows per CTA

// bias_cta[L][cta][128] <- the biases of the rows CTA `cta` owns in layer L
__global__ void mega_pack_bias_kernel(MegaWs* ws, int rq, int ro, int r1, int r2) {
  const int L = blockIdx.y, cta = blockIdx.x, t = threadIdx.x;  // 128 threads
  const ma_decoder_weights& W = ws->w;
  __half v = __float2half_rn(0.0f);
  if (t < BIAS_OUT) {
    const int n = cta * rq + t;
    if (t < rq && n < QKV) v = reinterpret_cast<const __half*>(W.bqkv[L])[n];
  } else if (t < BIAS_FC1) {
    const int i = t - BIAS_OUT, n = cta * ro + i;
    if (i < ro && n < HID) v = reinterpret_cast<const __half*>(W.bo[L])[n];
  } else if (t < BIAS_FC2) {
    const int i = t - BIAS_FC1, n = cta * r1 + i;
    if (i < r1 && n < FFN) v = reinterpret_cast<const __half*>(W.b1[L])[n];
  } else {
    const int i = t - BIAS_FC2, n = cta * r2 + i;
    if (i < r2 && n < HID) v = reinterpret_cast<const __half*>(W.b2[L])[n];
  }
  ws->bias_cta[((size_t)L * 160 + cta) * 128 + t] = v;
}

int mega_prepare(const ma_decoder_weights* w, void* mega_ws, cudaStream_t st) {
  MegaWs* ws = reinterpret_cast<MegaWs*>(mega_ws);
  if (cudaMemsetAsync(ws, 0, offsetof(MegaWs, bias_cta), st) != cudaSuccess) return 1;  // all epochs 0
  if (cudaMemcpyAsync(&ws->w, w, sizeof(ma_decoder_weights), cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
  const int rq = mega_rpc(QKV), ro = mega_rpc(HID), r1 = mega_rpc(FFN), r2 = mega_rpc(HID);
  if (rq > 32 || ro > 16 || r1 > 64 || r2 > 16) {
    set_error("mega: rows per CTA out of range");
    return 1;
  }
  mega_pack_bias_kernel<<<dim3(160, w->n_layers), 128, 0, st>>>(ws, rq, ro, r1, r2);
  count_launch();
  return check_launch("mega_pack_bias_kernel") ? 0 : 1;
}

int mega_error_flag_offset() { return (int)offsetof(MegaWs, error); }
int mega_trace_offset() { return (int)offsetof(MegaWs, trace); }

int mega_enqueue(const ma_decoder_weights* w, SeqState s, int tmax, __half* kv, void* mega_ws, const SampleArgs& sa,
                 int n_steps, int step_base, int trace, cudaStream_t st) {
  if (tmax > MAX_CHUNKS * MA_ATTN_CHUNK) {
    set_error("mega: tmax=%d exceeds %d keys", tmax, MAX_CHUNKS * MA_ATTN_CHUNK);
    return 1;
  }
  const int sms = mega_sms();
  auto rpc = [&](int N) { return mega_rpc(N); };
  MegaArgs a;
  memset(&a, 0, sizeof(a));
  a.ws = reinterpret_cast<MegaWs*>(mega_ws);
  a.s = s;
  a.kv = kv;
  a.T = tmax;
  a.n_steps = n_steps;
  a.step_base = step_base;
  a.max_new = sa.max_new; a.eos_id = sa.eos_id; a.pad_id = sa.pad_id;
  a.rows_qkv = rpc(QKV); a.rows_out = rpc(HID); a.rows_fc1 = rpc(FFN); a.rows_fc2 = rpc(HID); a.rows_lm = rpc(w->vocab);
  a.out_ids = sa.out_ids; a.forced = sa.forced; a.logits_out = sa.logits_out; a.all_done = sa.all_done;
  a.nkeys_next = sa.nkeys_next;
  a.trace = trace;
  if (a.rows_lm > a.rows_qkv + a.rows_out + a.rows_fc1 || a.rows_lm > 64 || a.rows_fc1 > 64) {
    set_error("mega: rows per CTA out of range (lm %d, fc1 %d)", a.rows_lm, a.rows_fc1);
    return 1;
  }
  // Every CTA must produce fc1 and lm_head rows (they are what orders buffer reuse), so the grid is the
  // number of CTAs that own fc1 rows; CTAs beyond the rows of a smaller matrix idle in that phase only.
  const int grid = (FFN + a.rows_fc1 - 1) / a.rows_fc1;
  if (grid > sms || (w->vocab + a.rows_lm - 1) / a.rows_lm != grid || grid < NHEAD) {
    set_error("mega: unsupported SM count %d (grid %d)", sms, grid);
    return 1;
  }
  const size_t smem = sizeof(MegaSmem) + ((size_t)(a.rows_qkv + a.rows_out + a.rows_fc1) * HID + (size_t)a.rows_fc2 * FFN) * 2;
  static size_t attr_set = 0;
  if (smem > attr_set) {
    if (cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("mega: cannot get %zu bytes of shared memory", smem);
      cudaGetLastError();
      return 1;
    }
    attr_set = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident (they wait on each other's data)
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, decode_mega_kernel, a);
  count_launch();
  return check_launch("decode_mega_kernel") ? 0 : 1;
}

}  // namespace ma
