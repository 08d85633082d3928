// decode_mega.cu -- batch-1 greedy decode as ONE persistent kernel: one CTA per SM runs every phase
// of every layer of up to `n_steps` tokens, separated by grid-wide barriers.
//
// Why: a token is 24 x (qkv, attention, out_proj, fc1, fc2) + lm_head = 121 dependent phases that
// together must stream 623.5 MB of weights (+ the KV cache) from HBM in ~100 us.  As separate
// kernels (decode_fast.cu) every phase pays a launch boundary (4.8 us measured even with PDL).  Here
//   * each CTA owns a fixed block of rows of every weight matrix (contiguous bytes), staged through
//     four shared-memory buffers (qkv 42 KB, out_proj 14 KB, fc1 56 KB, fc2 56 KB) that are refilled by a
//     single bulk async copy (TMA 1-D, mbarrier completion) as soon as the phase that read them ends,
//     i.e. the weights of layer L+1 are in flight while layer L computes -- HBM streams continuously;
//   * the K/V rows an SM needs for attention are prefetched into registers before the qkv phase;
//   * activations are exchanged through L2 (ld.global.cg) and the residual stream lives in shared
//     memory (every CTA recomputes the LayerNorms redundantly);
//   * a phase boundary is one release/acquire counter barrier (~0.5 us) instead of a kernel launch.
// Arithmetic is the canonical order of DESIGN.md section 3: results are bit-identical to
// gemm_canon.cu / attention.cu / decode_fast.cu and to the CPU oracle.
#include "canon.cuh"
#include "internal.h"

namespace ma {

constexpr int MG_THREADS = 256;
constexpr int MG_WARPS = 8;
constexpr int PARTF = 66;  // o[64], max, sum

struct MegaWs {
  __half q[HID];
  __half attn16[HID];
  __half y16[HID];
  __half f16[FFN];
  float cand_val[256];
  int cand_idx[256];
  unsigned int bars[2];    // grid barrier counters; launch i uses slot i & 1 and clears the other on exit
  int error;               // 1: barrier timeout
  int head_cnt[NHEAD];     // last-arriver counters of the attention merge
  unsigned long long trace[8 * 160];
  ma_decoder_weights w;    // device copy of the weight table
  alignas(256) float part[1];  // [NHEAD][max_chunks][66], sized by mega_workspace_bytes()
};

struct MegaArgs {
  MegaWs* ws;
  SeqState s;
  __half* kv;      // [layer][kv][head][T][64]   (batch 1)
  long T;
  int n_steps, max_new, eos_id, pad_id, max_chunks, bar_slot;
  int rows_qkv, rows_out, rows_fc1, rows_fc2, rows_lm;  // rows per CTA of each matrix
  int32_t* out_ids;
  const int32_t* forced;
  __half* logits_out;
  int* all_done;
  int* nkeys_next;
  int trace;
};

// ---- grid barrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct GridBar {
  unsigned int* ctr;
  unsigned int target;
  int* err;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      red_release_add(ctr, 1u);
      unsigned int spins = 0;
      while ((int)(ld_acquire(ctr) - target) < 0) {
        if (++spins > (1u << 27)) {  // ~ seconds: never hang the GPU
          *err = 1;
          break;
        }
      }
    }
    __syncthreads();
  }
};

// ---- shared memory layout ---------------------------------------------------------------------------
struct alignas(128) MegaSmem {
  uint64_t bar[4];        // full barriers of buffers D (qkv), C (out), A (fc1), B (fc2)
  float red[8];
  float wmax[8];
  float ared[8][65];
  float bval[8];
  int bidx[8];
  int last;
  alignas(16) float hres[HID];   // residual stream
  alignas(16) __half xs[FFN];    // fp16 input vector of the current GEMV
};

// y[n] = fp16(dot(W[n], x) + b[n]) for the rows of this CTA held in shared memory `sw` ([nrows][K]);
// `emit(n_global, fp16 value)` is called by lane 0.  Warp w takes rows w, w+8, ... two at a time.
template <int K, typename Emit>
__device__ __forceinline__ void gemv_rows(const __half* sw, int nrows, int row0, const __half* bias, const __half* xs,
                                          int warp, int lane, Emit emit) {
  constexpr int G = K / 256;
  uint4 xp[G];
#pragma unroll
  for (int g = 0; g < G; g++) xp[g] = *reinterpret_cast<const uint4*>(xs + 256 * g + 8 * lane);
  for (int r = warp; r < nrows; r += 2 * MG_WARPS) {
    const int r2 = r + MG_WARPS;
    const bool two = r2 < nrows;
    const __half* w0 = sw + (size_t)r * K + 8 * lane;
    const __half* w1 = sw + (size_t)(two ? r2 : r) * K + 8 * lane;
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int g = 0; g < G; g++) {
      float xf[8], f0[8], f1[8];
      unpack8(xp[g], xf);
      unpack8(*reinterpret_cast<const uint4*>(w0 + 256 * g), f0);
      unpack8(*reinterpret_cast<const uint4*>(w1 + 256 * g), f1);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        a0 = ffma(f0[j], xf[j], a0);
        a1 = ffma(f1[j], xf[j], a1);
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      const int n0 = row0 + r;
      emit(n0, __float2half_rn(fadd(a0, bias ? __half2float(bias[n0]) : 0.0f)));
      if (two) {
        const int n1 = row0 + r2;
        emit(n1, __float2half_rn(fadd(a1, bias ? __half2float(bias[n1]) : 0.0f)));
      }
    }
  }
}

__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// refill a weight buffer: rows [row0, row0+nrows) of W[N][K] -> smem (one bulk copy), thread 0 only
__device__ __forceinline__ void refill(__half* dst, const void* W, int N, int K, int rows_per_cta, uint64_t* bar) {
  const int row0 = blockIdx.x * rows_per_cta;
  const int nrows = max(0, min(rows_per_cta, N - row0));
  fence_proxy_async();
  if (nrows > 0) {
    const uint32_t bytes = (uint32_t)nrows * K * 2;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(dst, reinterpret_cast<const __half*>(W) + (size_t)row0 * K, bytes, bar);
  } else {
    mbar_expect_tx(bar, 0);  // plain arrival so that the phase still completes
  }
}

// rows [lo, hi) (relative to this CTA's first lm row) of lm_head -> dst (thread 0 only)
__device__ __forceinline__ void refill_lm(__half* dst, const void* lm, int row0_lm, int n_lm, int lo, int hi,
                                          uint64_t* bar) {
  hi = min(hi, n_lm);
  fence_proxy_async();
  if (hi > lo) {
    const uint32_t bytes = (uint32_t)(hi - lo) * HID * 2;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(dst, reinterpret_cast<const __half*>(lm) + (size_t)(row0_lm + lo) * HID, bytes, bar);
  } else {
    mbar_expect_tx(bar, 0);
  }
}

__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(MegaArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  MegaSmem& sm = *reinterpret_cast<MegaSmem*>(smem_raw);
  __half* bufD = reinterpret_cast<__half*>(smem_raw + sizeof(MegaSmem));   // qkv rows
  __half* bufC = bufD + (size_t)a.rows_qkv * HID;                          // out_proj rows
  __half* bufA = bufC + (size_t)a.rows_out * HID;                          // fc1 rows
  __half* bufB = bufA + (size_t)a.rows_fc1 * HID;                          // fc2 rows (K = 4096)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = lane >> 3, li = lane & 7;
  MegaWs* ws = a.ws;
  const ma_decoder_weights& W = ws->w;
  const int NL = W.n_layers;
  const long T = a.T;

  GridBar gb;
  gb.ctr = &ws->bars[a.bar_slot];   // zero at launch: cleared by the previous launch (or the host memset)
  gb.err = &ws->error;
  gb.target = 0;

  uint32_t parD = 0, parC = 0, parA = 0, parB = 0;
  if (tid == 0) {
    for (int i = 0; i < 4; i++) mbar_init(&sm.bar[i], 1);
    mbar_fence_init();
    refill(bufD, W.wqkv[0], QKV, HID, a.rows_qkv, &sm.bar[0]);
    refill(bufC, W.wo[0], HID, HID, a.rows_out, &sm.bar[1]);
    refill(bufA, W.w1[0], FFN, HID, a.rows_fc1, &sm.bar[2]);
    refill(bufB, W.w2[0], HID, FFN, a.rows_fc2, &sm.bar[3]);
  }
  __syncthreads();

  // generation state, identical in every CTA
  int pos = a.s.pos[0], gen = a.s.gen[0], tok = a.s.tok[0], fin = a.s.finished[0];
  unsigned long long* tr = (a.trace && blockIdx.x == 0 && tid == 0) ? ws->trace : nullptr;
  int tri = 0;

  const int row0_qkv = blockIdx.x * a.rows_qkv, n_qkv = max(0, min(a.rows_qkv, QKV - row0_qkv));
  const int row0_out = blockIdx.x * a.rows_out, n_out = max(0, min(a.rows_out, HID - row0_out));
  const int row0_fc1 = blockIdx.x * a.rows_fc1, n_fc1 = max(0, min(a.rows_fc1, FFN - row0_fc1));
  const int row0_fc2 = blockIdx.x * a.rows_fc2, n_fc2 = max(0, min(a.rows_fc2, HID - row0_fc2));
  const int row0_lm = blockIdx.x * a.rows_lm, n_lm = max(0, min(a.rows_lm, W.vocab - row0_lm));

  for (int step = 0; step < a.n_steps; step++) {
    if (gen >= a.max_new || fin) break;  // uniform across the grid
    const int nkeys = pos + 1;
    const int nch = (nkeys + MA_ATTN_CHUNK - 1) / MA_ATTN_CHUNK;
    const int nitems = nch * NHEAD;
    if (tr && tri < 150) tr[tri++] = gtimer();

    for (int L = 0; L < NL; L++) {
      __half* kc = a.kv + ((size_t)(L * 2 + 0)) * NHEAD * T * HD;
      __half* vc = a.kv + ((size_t)(L * 2 + 1)) * NHEAD * T * HD;

      // ---------------- K/V prefetch into registers: first attention item of this CTA (rows < pos are old)
      uint4 kreg[8], vreg[8];
      int item = blockIdx.x;
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

      // ---------------- qkv phase: input = token embedding (layer 0) or LN2 of the previous layer
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
          const uint2 u = __ldcg(reinterpret_cast<const uint2*>(ws->y16 + 4 * tid));
          const __half2* hh = reinterpret_cast<const __half2*>(&u);
          const float2 p0 = __half22float2(hh[0]), p1 = __half22float2(hh[1]);
          v[0] = fadd(hv.x, p0.x); v[1] = fadd(hv.y, p0.y); v[2] = fadd(hv.z, p1.x); v[3] = fadd(hv.w, p1.y);
          layernorm4(v, W.ln2g[L - 1], W.ln2b[L - 1], MA_LN_EPS, HID, sm.red);
        }
        *reinterpret_cast<float4*>(sm.hres + 4 * tid) = make_float4(v[0], v[1], v[2], v[3]);
        __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(sm.xs + 4 * tid) = u;
      }
      __syncthreads();
      mbar_wait(&sm.bar[0], parD);
      parD ^= 1;
      gemv_rows<HID>(bufD, n_qkv, row0_qkv, (const __half*)W.bqkv[L], sm.xs, warp, lane, [&](int n, __half hv) {
        if (n < HID) {
          ws->q[n] = hv;
        } else {
          const int e = (n - HID) & (HID - 1), head = e >> 6, d = e & 63;
          __half* c = (n < 2 * HID) ? kc : vc;
          c[((long)head * T + pos) * HD + d] = hv;
        }
      });
      __syncthreads();
      if (tid == 0) {
        if (L + 1 < NL) refill(bufD, W.wqkv[L + 1], QKV, HID, a.rows_qkv, &sm.bar[0]);
        else refill_lm(bufD, W.lm_head, row0_lm, n_lm, 0, a.rows_qkv, &sm.bar[0]);  // lm rows live in D|C|A (contiguous)
      }
      gb.sync();
      if (tr && tri < 150) tr[tri++] = gtimer();

      // ---------------- attention phase: items (chunk c, head h) = blockIdx.x, + gridDim.x, ...
      for (int it = 0; item < nitems; item += gridDim.x, it++) {
        const int c = item >> 4, h = item & 15;
        const int len = min(MA_ATTN_CHUNK, nkeys - c * MA_ATTN_CHUNK);
        const long base = ((long)h * T + (long)c * MA_ATTN_CHUNK) * HD;
#pragma unroll
        for (int rho = 0; rho < 8; rho++) {
          const int r = 32 * rho + 4 * warp + grp;
          const int ap = c * MA_ATTN_CHUNK + r;
          // rows not prefetched: every row of a later item, and the current token's row (written this phase)
          if (r < len && (it > 0 || ap >= pos)) {
            kreg[rho] = ldcg16(kc + base + (long)r * HD + 8 * li);
            vreg[rho] = ldcg16(vc + base + (long)r * HD + 8 * li);
          }
        }
        float qf[8];
        unpack8(ldcg16(ws->q + h * HD + 8 * li), qf);
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
        __syncthreads();  // previous item's readers of wmax / ared are done
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
        float* part = ws->part + ((long)h * a.max_chunks) * PARTF;
        float rsum = 0.0f;
        if (tid < 65) {
          float x[8];
#pragma unroll
          for (int w2 = 0; w2 < 8; w2++) x[w2] = sm.ared[w2][tid];
          rsum = fadd(fadd(fadd(x[0], x[1]), fadd(x[2], x[3])), fadd(fadd(x[4], x[5]), fadd(x[6], x[7])));
        }
        if (nch == 1) {
          __syncthreads();
          if (tid < 65) sm.ared[0][tid] = rsum;
          __syncthreads();
          if (tid < 64) ws->attn16[h * HD + tid] = __float2half_rn(__fdiv_rn(sm.ared[0][tid], sm.ared[0][64]));
        } else {
          if (tid < 65) {
            part[c * PARTF + (tid < 64 ? tid : 65)] = rsum;
            if (tid == 64) part[c * PARTF + 64] = cmax;
          }
          __threadfence();
          __syncthreads();
          if (tid == 0) {
            const int prev = atomicAdd(&ws->head_cnt[h], 1);
            sm.last = (prev == nch - 1);
            if (sm.last) ws->head_cnt[h] = 0;
          }
          __syncthreads();
          if (sm.last) {
            __threadfence();
            if (tid < 64) {
              float M = -INFINITY;
              for (int cc = 0; cc < nch; cc++) M = fmaxf(M, __ldcg(part + cc * PARTF + 64));
              float Lsum = 0.0f, O = 0.0f;
              for (int cc = 0; cc < nch; cc++) {
                const float wgt = ma_exp(fsub(__ldcg(part + cc * PARTF + 64), M));
                Lsum = ffma(__ldcg(part + cc * PARTF + 65), wgt, Lsum);
                O = ffma(__ldcg(part + cc * PARTF + tid), wgt, O);
              }
              ws->attn16[h * HD + tid] = __float2half_rn(__fdiv_rn(O, Lsum));
            }
          }
        }
      }
      gb.sync();
      if (tr && tri < 150) tr[tri++] = gtimer();

      // ---------------- out_proj phase
      if (tid < HID / 8) *reinterpret_cast<uint4*>(sm.xs + 8 * tid) = ldcg16(ws->attn16 + 8 * tid);
      __syncthreads();
      mbar_wait(&sm.bar[1], parC);
      parC ^= 1;
      gemv_rows<HID>(bufC, n_out, row0_out, (const __half*)W.bo[L], sm.xs, warp, lane,
                     [&](int n, __half hv) { ws->y16[n] = hv; });
      __syncthreads();
      if (tid == 0) {
        if (L + 1 < NL) refill(bufC, W.wo[L + 1], HID, HID, a.rows_out, &sm.bar[1]);
        else refill_lm(bufC, W.lm_head, row0_lm, n_lm, a.rows_qkv, a.rows_qkv + a.rows_out, &sm.bar[1]);
      }
      gb.sync();
      if (tr && tri < 150) tr[tri++] = gtimer();

      // ---------------- fc1 phase: input = LN1(hres + out_proj)
      {
        const float4 hv = *reinterpret_cast<const float4*>(sm.hres + 4 * tid);
        const uint2 u = __ldcg(reinterpret_cast<const uint2*>(ws->y16 + 4 * tid));
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
        const float2 p0 = __half22float2(hh[0]), p1 = __half22float2(hh[1]);
        float v[4] = {fadd(hv.x, p0.x), fadd(hv.y, p0.y), fadd(hv.z, p1.x), fadd(hv.w, p1.y)};
        layernorm4(v, W.ln1g[L], W.ln1b[L], MA_LN_EPS, HID, sm.red);
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
      gemv_rows<HID>(bufA, n_fc1, row0_fc1, (const __half*)W.b1[L], sm.xs, warp, lane, [&](int n, __half hv) {
        if (__half2float(hv) < 0.0f) hv = __float2half_rn(0.0f);
        ws->f16[n] = hv;
      });
      __syncthreads();
      if (tid == 0) {
        if (L + 1 < NL) refill(bufA, W.w1[L + 1], FFN, HID, a.rows_fc1, &sm.bar[2]);
        else refill_lm(bufA, W.lm_head, row0_lm, n_lm, a.rows_qkv + a.rows_out, a.rows_lm, &sm.bar[2]);
      }
      gb.sync();
      if (tr && tri < 150) tr[tri++] = gtimer();

      // ---------------- fc2 phase
      for (int i = tid; i < FFN / 8; i += MG_THREADS) *reinterpret_cast<uint4*>(sm.xs + 8 * i) = ldcg16(ws->f16 + 8 * i);
      __syncthreads();
      mbar_wait(&sm.bar[3], parB);
      parB ^= 1;
      gemv_rows<FFN>(bufB, n_fc2, row0_fc2, (const __half*)W.b2[L], sm.xs, warp, lane,
                     [&](int n, __half hv) { ws->y16[n] = hv; });
      __syncthreads();
      if (tid == 0) refill(bufB, W.w2[(L + 1 < NL) ? L + 1 : 0], HID, FFN, a.rows_fc2, &sm.bar[3]);
      gb.sync();
      if (tr && tri < 150) tr[tri++] = gtimer();
    }

    // ---------------- lm_head on LN2 of the last layer + greedy pick
    {
      const float4 hv = *reinterpret_cast<const float4*>(sm.hres + 4 * tid);
      const uint2 u = __ldcg(reinterpret_cast<const uint2*>(ws->y16 + 4 * tid));
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
      const float2 p0 = __half22float2(hh[0]), p1 = __half22float2(hh[1]);
      float v[4] = {fadd(hv.x, p0.x), fadd(hv.y, p0.y), fadd(hv.z, p1.x), fadd(hv.w, p1.y)};
      layernorm4(v, W.ln2g[NL - 1], W.ln2b[NL - 1], MA_LN_EPS, HID, sm.red);
      __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
      uint2 uo;
      uo.x = *reinterpret_cast<uint32_t*>(&h0);
      uo.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(sm.xs + 4 * tid) = uo;
    }
    __syncthreads();
    // the lm rows of this CTA sit contiguously in D|C|A (refilled after the last layer's qkv / out_proj / fc1)
    mbar_wait(&sm.bar[0], parD);
    mbar_wait(&sm.bar[1], parC);
    mbar_wait(&sm.bar[2], parA);
    parD ^= 1; parC ^= 1; parA ^= 1;
    float bestv = -INFINITY;
    int besti = 0x7fffffff;
    gemv_rows<HID>(bufD, n_lm, row0_lm, nullptr, sm.xs, warp, lane, [&](int n, __half hv) {
      if (a.logits_out) a.logits_out[(long)gen * W.vocab + n] = hv;
      const float v = __half2float(hv);
      if (v > bestv || (v == bestv && n < besti)) { bestv = v; besti = n; }
    });
    if (lane == 0) { sm.bval[warp] = bestv; sm.bidx[warp] = besti; }
    __syncthreads();
    if (tid == 0) {
      float bv = sm.bval[0];
      int bi = sm.bidx[0];
      for (int w2 = 1; w2 < MG_WARPS; w2++)
        if (sm.bval[w2] > bv || (sm.bval[w2] == bv && sm.bidx[w2] < bi)) { bv = sm.bval[w2]; bi = sm.bidx[w2]; }
      ws->cand_val[blockIdx.x] = bv;
      ws->cand_idx[blockIdx.x] = bi;
      // weights of the next token's first layer
      refill(bufD, W.wqkv[0], QKV, HID, a.rows_qkv, &sm.bar[0]);
      refill(bufC, W.wo[0], HID, HID, a.rows_out, &sm.bar[1]);
      refill(bufA, W.w1[0], FFN, HID, a.rows_fc1, &sm.bar[2]);
    }
    gb.sync();
    {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int i = tid; i < (int)gridDim.x; i += MG_THREADS) {
        const float v = __ldcg(&ws->cand_val[i]);
        const int ix = __ldcg(&ws->cand_idx[i]);
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
      if (blockIdx.x == 0 && tid == 0) {
        if (gen < a.max_new) a.out_ids[gen] = ntok;
        if (!fin) a.s.lens[0] = gen + 1;
      }
      if (!fin && ntok == a.eos_id) fin = 1;
      tok = ntok;
      gen += 1;
      pos += 1;
    }
    if (tr && tri < 150) tr[tri++] = gtimer();
  }

  // every buffer has a refill in flight here: drain them before the shared memory is released
  mbar_wait(&sm.bar[0], parD);
  mbar_wait(&sm.bar[1], parC);
  mbar_wait(&sm.bar[2], parA);
  mbar_wait(&sm.bar[3], parB);
  if (blockIdx.x == 0 && tid == 0) {
    ws->bars[a.bar_slot ^ 1] = 0;
    a.s.pos[0] = pos; a.s.gen[0] = gen; a.s.tok[0] = tok; a.s.finished[0] = fin;
    if (a.nkeys_next) *a.nkeys_next = pos + 1;
    if (a.all_done) *a.all_done = fin;
  }
}

// ---- host side -----------------------------------------------------------------------------------
static int g_mega_sms = 0;

size_t mega_workspace_bytes() { return sizeof(MegaWs) + (size_t)NHEAD * 72 * PARTF * sizeof(float) + 256; }

int mega_prepare(const ma_decoder_weights* w, void* mega_ws, cudaStream_t st) {
  MegaWs* ws = reinterpret_cast<MegaWs*>(mega_ws);
  if (cudaMemsetAsync(ws, 0, offsetof(MegaWs, w), st) != cudaSuccess) return 1;
  if (cudaMemcpyAsync(&ws->w, w, sizeof(ma_decoder_weights), cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
  return 0;
}

int mega_error_flag_offset() { return (int)offsetof(MegaWs, error); }
int mega_trace_offset() { return (int)offsetof(MegaWs, trace); }

int mega_enqueue(const ma_decoder_weights* w, SeqState s, int tmax, __half* kv, void* mega_ws, const SampleArgs& sa,
                 int n_steps, int bar_slot, int trace, cudaStream_t st) {
  if (!g_mega_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_mega_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_mega_sms <= 0 || g_mega_sms > 256) g_mega_sms = 148;
  }
  const int grid = g_mega_sms;
  auto rpc = [&](int N) { return (N + grid - 1) / grid; };
  MegaArgs a;
  memset(&a, 0, sizeof(a));
  a.ws = reinterpret_cast<MegaWs*>(mega_ws);
  a.s = s;
  a.kv = kv;
  a.T = tmax;
  a.n_steps = n_steps;
  a.max_new = sa.max_new; a.eos_id = sa.eos_id; a.pad_id = sa.pad_id;
  a.max_chunks = 72;
  a.bar_slot = bar_slot & 1;
  a.rows_qkv = rpc(QKV); a.rows_out = rpc(HID); a.rows_fc1 = rpc(FFN); a.rows_fc2 = rpc(HID); a.rows_lm = rpc(w->vocab);
  a.out_ids = sa.out_ids; a.forced = sa.forced; a.logits_out = sa.logits_out; a.all_done = sa.all_done;
  a.nkeys_next = sa.nkeys_next;
  a.trace = trace;
  if (a.rows_lm > a.rows_qkv + a.rows_out + a.rows_fc1) {
    set_error("mega: lm_head rows per CTA (%d) exceed the qkv+out+fc1 buffers", a.rows_lm);
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
  at[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident (the kernel spins on a grid barrier)
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, decode_mega_kernel, a);
  count_launch();
  return check_launch("decode_mega_kernel") ? 0 : 1;
}

}  // namespace ma
