// decode_mega.cu -- batch-1 greedy decode as ONE persistent kernel: 144 CTAs (one per SM) = 16 groups x 9 CTAs run
// every phase of every layer of up to `n_steps` tokens.
//
// A token is 24 x (qkv, attention, out_proj, LN, fc1, fc2, LN) + lm_head dependent phases that together stream
// 621 MB of weights (+ the KV cache) in ~100 us of HBM time, i.e. ~4 us per layer: the step is bound by the LATENCY
// of the hand-offs between phases, not by bandwidth.  Round 1 ran six chip-wide all-gathers per layer (one of them
// 8 KB wide); this version is partitioned by attention head so that most hand-offs stay inside a group of 9 CTAs
// and only two reductions per layer cross the chip:
//   * group g owns attention head g: its 9 CTAs compute the 192 q/k/v rows of the head (20 or 22 each), exchange
//     them inside the group, split the KV chunks of the head among their 18 attention teams; every CTA of the group
//     reads all chunk partials and merges them redundantly (no second hand-off for the merged head output);
//   * out_proj is split-K by head: group g multiplies its head's 64 attention outputs with the 64-column slice of
//     W_o (CTA j: rows 1024j/9 .. 1024(j+1)/9) -> 16 partial vectors; 128 reducer CTAs (8 rows each) add the 16
//     partials of their rows (balanced tree), add the bias, round to fp16 and publish 4 words; every CTA gathers the
//     512 words for its redundant residual + LayerNorm;
//   * group g also owns fc1 rows 256g..256g+255 (28 or 30 per CTA); the 256 activations are exchanged inside the
//     group and fc2 is split-K by group the same way (CTA j: its rows of the 256-column slice of W_2) -> the second
//     chip-wide reduction of the layer.
// Hardware thread-block clusters were measured first (tools/microbench_cluster.cu, profiles/microbench_cluster_r02.txt):
// only 15 clusters of 8 CTAs of this shape are co-resident on the B200 (GPC floor-sweeping), 16 are needed, and a
// DSMEM st.async all-gather (0.26-0.62 us) is no cheaper than a flagged-word exchange among 9 CTAs through L2.  So
// every hand-off is the round-1 mechanism: {payload, epoch} 8-byte words written with st.volatile and polled with
// ld.volatile (8-byte stores are single-copy atomic: a reader that sees the epoch sees the data -- NCCL "LL").
// The segmented accumulation order that split-K implies (16 segment dots + balanced tree) is the canonical order of
// out_proj / fc2: gemm_canon.cu, decode_fast.cu and the CPU oracle implement the same order,
// so token ids stay bit-identical across all paths.
//   * every weight slice of a CTA is one contiguous byte range (W_o / W_2 are repacked per head / per 256 columns once
//     per generate) pulled by cp.async.bulk into four shared-memory buffers (qkv 44 KB, out 14 KB, fc1 60 KB, fc2
//     57 KB) that are refilled by the last warp that leaves the phase: layer L+1 streams in while layer L computes;
//   * K/V of a team's first chunk are prefetched into registers before the qkv phase; the chunks of the next layer
//     are prefetched into L2 (cp.async.bulk.prefetch.L2) right after the attention phase of this one.
// Every wait is bounded (globaltimer deadline): a time-out sets MegaWs::error, the kernel stops emitting tokens and
// reports lens = -1, which the callers of ma_decode_generate turn into an error (no silently wrong mesh).
#include <stdlib.h>

#include "canon.cuh"
#include "internal.h"

namespace ma {

constexpr int MG_THREADS = 512;  // 16 warps
constexpr int MG_WARPS = 16;
constexpr int TEAM = 256;        // threads of one attention team (8 warps = 32 group-lanes, the canonical structure)
constexpr int GS = 9;            // CTAs per group
constexpr int NG = 16;           // groups = attention heads = 256-wide slices of the fc1 activations
constexpr int MG_GRID = GS * NG; // 144
constexpr int NRED = 128;        // reducer CTAs: CTA i < 128 owns rows 8i..8i+7 of the out_proj / fc2 vectors
constexpr int PARTF = 66;        // o[64], max, sum
constexpr int MAX_CHUNKS = 60;   // 15360 keys (config 5 needs 58)

// rows per CTA (maxima over the ranks j of a group)
constexpr int ROWS_QKV = 22;     // 10 or 11 row pairs of the head's 96 (q | k | v)
constexpr int ROWS_OUT = 114;    // 113 or 114 rows of the head's 64-column slice of W_o
constexpr int ROWS_FC1 = 30;     // 14 or 15 row pairs of the group's 128
constexpr int ROWS_FC2 = 114;    // 113 or 114 rows of the group's 256-column slice of W_2
constexpr int ROWS_LM = 58;      // 58 * 144 >= 8256
constexpr int BYTES_D = ROWS_QKV * HID * 2, BYTES_C = ROWS_OUT * HD * 2, BYTES_A = ROWS_FC1 * HID * 2,
              BYTES_B = ROWS_FC2 * 256 * 2;
static_assert(BYTES_D + BYTES_C + BYTES_A >= ROWS_LM * HID * 2, "the lm_head rows of a CTA live in D|C|A");
static_assert(BYTES_D % 128 == 0 && BYTES_C % 128 == 0 && BYTES_A % 128 == 0 && BYTES_B % 128 == 0, "alignment");

// packed per-CTA biases of a layer (fp16): qkv rows, fc1 rows, then (reducers) the 8 out_proj and 8 fc2 rows
constexpr int BIAS_QKV = 0, BIAS_FC1 = 22, BIAS_OUT = 52, BIAS_FC2 = 60, BIAS_N = 72;

__host__ __device__ constexpr int qkv_pair0(int j) { return (32 * j) / 3; }    // of the head's 96 word pairs
__host__ __device__ constexpr int fc1_pair0(int j) { return (128 * j) / 9; }   // of the group's 128 word pairs
__host__ __device__ constexpr int slice_row0(int j) { return (1024 * j) / 9; } // of the 1024 out_proj / fc2 rows

struct MegaWs {
  uint2 qkv_w[NG * 96];             // per group: q | k | v of the current token, {2 x fp16, epoch}
  uint2 f_w[NG * 128];              // per group: its 256 fc1 activations
  uint2 pa_w[HID * NG];             // out_proj partials [row][head]: {fp32 bits, epoch}
  uint2 pb_w[HID * NG];             // fc2 partials      [row][group]
  uint2 ya_w[HID / 2];              // reduced out_proj vector (+bias), {2 x fp16, epoch}
  uint2 yb_w[HID / 2];              // reduced fc2 vector
  uint2 cand_w[MG_GRID * 2];        // {value bits, epoch}, {index, epoch}
  uint2 part_w[NG * MAX_CHUNKS * PARTF];  // attention chunk partials {fp32 bits, epoch}
  int error;                        // != 0: a wait timed out (code in the low byte, CTA above it)
  int pad_[3];
  unsigned long long trace[1280];
  unsigned long long trace_cta[MG_GRID * 16];  // per-CTA stamps of one (step, layer): skew analysis
  int fail[8];                     // first time-out: {code, cta, tid, epoch waited for, epoch seen, word offset in ws, -, -}
  int wprog[MG_GRID * 16];         // debug (trace & 2): epoch of the last out_proj partial each warp published
  alignas(16) int where[MG_GRID * 4];   // per CTA: last phase reached {step, layer, phase, -} (post-mortem of a time-out)
  ma_decoder_weights w;             // device copy of the weight table
  alignas(256) __half bias_cta[MA_MAX_LAYERS * MG_GRID * BIAS_N];
  alignas(256) __half wo_p[(size_t)MA_MAX_LAYERS * HID * HID];   // [layer][head][row][64]
  alignas(256) __half w2_p[(size_t)MA_MAX_LAYERS * HID * FFN];   // [layer][group][row][256]
};

struct MegaArgs {
  MegaWs* ws;
  SeqState s;
  __half* kv;  // [layer][kv][head][T][64]   (batch 1)
  long T;
  int n_steps, step_base, max_new, eos_id, pad_id;
  int32_t* out_ids;
  const int32_t* forced;
  __half* logits_out;
  int* all_done;
  int* nkeys_next;
  int trace;
  int fault;                      // test hook: CTA `fault - 1` withholds its out_proj partials from step 2 on
  unsigned long long timeout_ns;  // bound of every wait
  unsigned backoff_ns;
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---- flagged-word exchange through L2 -----------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint2* p, uint32_t data, uint32_t ep) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(data), "r"(ep) : "memory");
}
__device__ __forceinline__ uint2 ll_load1(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ll_load2(const uint2* p) {  // two words
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile_i32(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// bounded waits: the fast path is plain polling; every 256 polls the slow path looks at the global error flag and at
// a globaltimer deadline
struct WaitCtx {
  int* err;
  unsigned long long timeout_ns;
  int cta;
  unsigned backoff_ns;   // experiment: sleep between unsuccessful polls (0 = spin)
  int* fail;
  const void* base;
};
enum { ERR_WBAR = 1, ERR_QKV = 2, ERR_PART = 3, ERR_RED = 4, ERR_Y = 5, ERR_F = 6, ERR_CAND = 7 };
__device__ __noinline__ bool wait_slow(const WaitCtx& wc, unsigned long long& t0, int code, const void* addr = nullptr,
                                       uint32_t ep = 0, uint32_t seen = 0) {
  if (ld_volatile_i32(wc.err)) return true;
  const unsigned long long now = gtimer();
  if (t0 == 0) { t0 = now; return false; }
  if (now - t0 > wc.timeout_ns) {
    if (atomicCAS(wc.err, 0, code | (wc.cta << 8)) == 0) {
      wc.fail[0] = code; wc.fail[1] = wc.cta; wc.fail[2] = threadIdx.x; wc.fail[3] = (int)ep; wc.fail[4] = (int)seen;
      wc.fail[5] = addr ? (int)((const char*)addr - (const char*)wc.base) : -1;
    }
    return true;
  }
  return false;
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity, const WaitCtx& wc) {
  if (mbar_try(bar, parity)) return;
  unsigned long long t0 = 0;
  unsigned n = 0;
  while (!mbar_try(bar, parity)) {
    if ((++n & 15u) == 0 && wait_slow(wc, t0, ERR_WBAR)) return;
  }
}
// wait for N single words at p + i*stride (in words); returns the payloads
template <int N>
__device__ __forceinline__ void ll_wait_words(const uint2* p, long stride, uint32_t ep, uint32_t* out, const WaitCtx& wc,
                                              int code) {
  uint2 w[N];
#pragma unroll
  for (int i = 0; i < N; i++) w[i] = ll_load1(p + i * stride);
  unsigned n = 0;
  unsigned long long t0 = 0;
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (w[i].y != ep) {
        ok = false;
        w[i] = ll_load1(p + i * stride);
      }
    }
    if (ok) break;
    if (wc.backoff_ns) __nanosleep(wc.backoff_ns);
    if ((++n & 255u) == 0 && wait_slow(wc, t0, code, p, ep, w[0].y)) break;
  }
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = w[i].x;
}
// wait for N consecutive 16-byte units (2 words each) starting at p: all loads are in flight before any flag is
// checked and only the units that are not there yet are polled again; out[i] = the two payloads of unit i
template <int N>
__device__ __forceinline__ void ll_wait_units(const uint2* p, uint32_t ep, uint2* out, const WaitCtx& wc, int code) {
  uint4 v[N];
#pragma unroll
  for (int i = 0; i < N; i++) v[i] = ll_load2(p + 2 * i);
  unsigned n = 0;
  unsigned long long t0 = 0;
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (v[i].y != ep || v[i].w != ep) {
        ok = false;
        v[i] = ll_load2(p + 2 * i);
      }
    }
    if (ok) break;
    if (wc.backoff_ns) __nanosleep(wc.backoff_ns);
    if ((++n & 255u) == 0 && wait_slow(wc, t0, code, p, ep, v[0].y != ep ? v[0].y : v[0].w)) break;
  }
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = make_uint2(v[i].x, v[i].z);
}

// ---- shared memory layout ---------------------------------------------------------------------------------------
struct alignas(128) MegaSmem {
  uint64_t wbar[4];   // full barriers of the weight buffers D (qkv), C (out_proj), A (fc1), B (fc2)
  uint64_t lnbar[2];  // ln1 / ln2 parameter regions
  uint64_t bbar[2];   // packed-bias double buffer
  unsigned int cnt[4];
  int errflag;
  // this CTA's share of every matrix (kept here rather than in registers: the attention phase needs the registers)
  int qp0, n_qp, qsec, qrow0;   // q/k/v word pairs [qp0, qp0 + n_qp) of the head's 96; section 0: q, 1: k, 2: v
  int fp0, n_fp;                // fc1 word pairs of the group's 128
  int sr0, n_sr;                // out_proj / fc2 slice rows of the 1024
  int row0_lm, n_lm, lm_bytes;
  float red[2][8];
  float wmax[2][8];
  float bval[MG_WARPS];
  int bidx[MG_WARPS];
  float mw[MAX_CHUNKS + 4];            // merge weights exp(m_c - M)
  const void* ptab[MA_MAX_LAYERS][6];  // wqkv, w1, ln1g, ln1b, ln2g, ln2b (kept on chip: a miss is ~1 us)
  const void* gtab[6];                 // lm_head, tok_table, extra, tok_pos, cond, pos
  alignas(16) float ln1[2 * HID];      // gamma | beta of self_attn_layer_norm of the current layer
  alignas(16) float ln2[2 * HID];      // gamma | beta of final_layer_norm
  alignas(16) __half bias[2][BIAS_N];
  alignas(16) __half xs[HID];          // fp16 input vector of the qkv / fc1 / lm_head GEMVs
  alignas(16) __half ab[HD];           // merged attention output of this head
  alignas(16) __half stage16[64];
  float ared[2][8][65];                // attention team scratch
  alignas(16) float part[MAX_CHUNKS * PARTF];  // chunk partials of this head (staged for the merge)
};

__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(TEAM) : "memory"); }
__device__ __forceinline__ void half_sync() { asm volatile("bar.sync 3, 256;" ::: "memory"); }   // warps 0-7
__device__ __forceinline__ void merge_sync() { asm volatile("bar.sync 4, 64;" ::: "memory"); }    // warps 0-1

__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// the last of `nwarps` warps to call this (after its last read of a weight buffer) gets true in lane 0
__device__ __forceinline__ bool last_warp_out(unsigned int* cnt, unsigned nwarps, int lane) {
  __syncwarp();
  bool last = false;
  if (lane == 0) {
    unsigned old;
    // relaxed: the shared-memory reads of this warp were issued (in order) before this atomic, and an acq_rel atomic
    // costs a MEMBAR that also waits for the warp's outstanding GLOBAL stores (the flagged words) -- ~1 us per phase
    asm volatile("atom.relaxed.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
    last = (old == nwarps - 1);
    if (last) *reinterpret_cast<volatile unsigned int*>(cnt) = 0;
  }
  return last;
}

// LayerNorm of 1024 values by the first 256 threads (thread t owns 4t..4t+3: the canonical block sum)
__device__ __forceinline__ void layernorm_1024(float* v, const float* gamma, const float* beta, float (*red)[8], int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const float inv = __fdiv_rn(1.0f, 1024.0f);
  float p = fadd(fadd(v[0], v[1]), fadd(v[2], v[3]));
  p = warp_sum(p);
  if (lane == 0) red[0][warp] = p;
  half_sync();
  const float mean = fmul(warp_tree(red[0], 8), inv);
  const float d0 = fsub(v[0], mean), d1 = fsub(v[1], mean), d2 = fsub(v[2], mean), d3 = fsub(v[3], mean);
  float q = fadd(fadd(fmul(d0, d0), fmul(d1, d1)), fadd(fmul(d2, d2), fmul(d3, d3)));
  q = warp_sum(q);
  if (lane == 0) red[1][warp] = q;
  half_sync();
  const float var = fmul(warp_tree(red[1], 8), inv);
  const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(fadd(var, MA_LN_EPS)));
  const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * tid);
  const float4 b = *reinterpret_cast<const float4*>(beta + 4 * tid);
  v[0] = ffma(fmul(d0, rstd), g.x, b.x);
  v[1] = ffma(fmul(d1, rstd), g.y, b.y);
  v[2] = ffma(fmul(d2, rstd), g.z, b.z);
  v[3] = ffma(fmul(d3, rstd), g.w, b.w);
}

// canonical K = 1024 dot products of NR rows that share x: lane l accumulates k = 256g + 8l + j (g major, j minor)
template <int NR>
__device__ __forceinline__ void gemv_k1024(const __half* const* w, const __half* xs, int lane, float* acc) {
#pragma unroll
  for (int i = 0; i < NR; i++) acc[i] = 0.0f;
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const uint4 xr = *reinterpret_cast<const uint4*>(xs + 256 * g + 8 * lane);
#pragma unroll
    for (int i = 0; i < NR; i++) acc[i] = dot8(*reinterpret_cast<const uint4*>(w[i] + 256 * g + 8 * lane), xr, acc[i]);
  }
}

__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(MegaArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __half* bufD = reinterpret_cast<__half*>(smem_raw);                                // qkv rows (K = 1024)
  __half* bufC = reinterpret_cast<__half*>(smem_raw + BYTES_D);                      // rows x 64 slice of W_o
  __half* bufA = reinterpret_cast<__half*>(smem_raw + BYTES_D + BYTES_C);            // fc1 rows
  __half* bufB = reinterpret_cast<__half*>(smem_raw + BYTES_D + BYTES_C + BYTES_A);  // rows x 256 slice of W_2
  MegaSmem& sm = *reinterpret_cast<MegaSmem*>(smem_raw + BYTES_D + BYTES_C + BYTES_A + BYTES_B);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int team = warp >> 3, wt = warp & 7, tl = tid & (TEAM - 1);   // attention team / warp and thread in it
  const int grp = lane >> 3, li = lane & 7;
  MegaWs* ws = a.ws;
  const int cta = blockIdx.x;
  const int g = cta / GS;   // group = head = fc1 slice
  const int j = cta % GS;   // rank in the group
  const WaitCtx wc = {&ws->error, a.timeout_ns, cta, a.backoff_ns, ws->fail, ws};

  const int NL = ws->w.n_layers, vocab = ws->w.vocab;
  for (int i = tid; i < NL * 6; i += MG_THREADS) {
    const int L = i / 6, k = i % 6;
    const ma_decoder_weights& W = ws->w;
    sm.ptab[L][k] = k == 0 ? W.wqkv[L] : k == 1 ? W.w1[L] : k == 2 ? (const void*)W.ln1g[L] : k == 3 ? (const void*)W.ln1b[L]
                  : k == 4 ? (const void*)W.ln2g[L] : (const void*)W.ln2b[L];
  }
  if (tid == 0) {
    const ma_decoder_weights& W = ws->w;
    sm.gtab[0] = W.lm_head; sm.gtab[1] = W.tok_table; sm.gtab[2] = W.extra; sm.gtab[3] = W.tok_pos; sm.gtab[4] = W.cond;
    sm.gtab[5] = W.pos;
    for (int i = 0; i < 4; i++) { mbar_init(&sm.wbar[i], 1); sm.cnt[i] = 0; }
    for (int i = 0; i < 2; i++) { mbar_init(&sm.lnbar[i], 1); mbar_init(&sm.bbar[i], 1); }
    mbar_fence_init();
    sm.errflag = ld_volatile_i32(&ws->error);
    sm.qp0 = qkv_pair0(j); sm.n_qp = qkv_pair0(j + 1) - sm.qp0;            // 10 or 11 pairs
    sm.qsec = sm.qp0 >> 5;                                               // a CTA never straddles q / k / v
    sm.qrow0 = sm.qsec * HID + g * HD + 2 * (sm.qp0 & 31);               // first row in the stacked [3072][1024] matrix
    sm.fp0 = fc1_pair0(j); sm.n_fp = fc1_pair0(j + 1) - sm.fp0;            // 14 or 15 pairs
    sm.sr0 = slice_row0(j); sm.n_sr = slice_row0(j + 1) - sm.sr0;          // 113 or 114 rows
    sm.row0_lm = cta * ROWS_LM; sm.n_lm = max(0, min(ROWS_LM, vocab - sm.row0_lm));
    sm.lm_bytes = sm.n_lm * HID * 2;
  }
  __syncthreads();
  if (sm.errflag) return;   // an earlier launch of this generate failed: nothing more is emitted (CTA-uniform)
  const long T = a.T;
  const bool reducer = cta < NRED;

  // ---- weight / parameter fills (one thread each) ----
  auto fill = [&](void* dst, const void* src, int bytes, uint64_t* bar) {   // bytes >= 0, multiple of 16
    fence_proxy_async();
    mbar_expect_tx(bar, (uint32_t)bytes);
    if (bytes > 0) bulk_g2s(dst, src, (uint32_t)bytes, bar);
  };
  auto fill_D = [&](int L) {
    fill(bufD, reinterpret_cast<const __half*>(sm.ptab[L][0]) + (size_t)sm.qrow0 * HID, 2 * sm.n_qp * HID * 2, &sm.wbar[0]);
  };
  auto fill_C = [&](int L) {
    fill(bufC, ws->wo_p + (size_t)L * HID * HID + ((size_t)g * HID + sm.sr0) * HD, sm.n_sr * HD * 2, &sm.wbar[1]);
  };
  auto fill_A = [&](int L) {
    fill(bufA, reinterpret_cast<const __half*>(sm.ptab[L][1]) + ((size_t)g * 256 + 2 * sm.fp0) * HID, 2 * sm.n_fp * HID * 2, &sm.wbar[2]);
  };
  auto fill_B = [&](int L) {
    fill(bufB, ws->w2_p + (size_t)L * HID * FFN + ((size_t)g * HID + sm.sr0) * 256, sm.n_sr * 256 * 2, &sm.wbar[3]);
  };
  auto fill_lm = [&](int which) {   // byte ranges of this CTA's lm_head rows as the buffers D, C, A become free
    const unsigned char* lm = reinterpret_cast<const unsigned char*>(sm.gtab[0]) + (size_t)sm.row0_lm * HID * 2;
    const int lo = which == 0 ? 0 : which == 1 ? BYTES_D : BYTES_D + BYTES_C;
    const int hi = which == 0 ? BYTES_D : which == 1 ? BYTES_D + BYTES_C : BYTES_D + BYTES_C + BYTES_A;
    fill(smem_raw + lo, lm + lo, max(0, min(hi, sm.lm_bytes) - lo), &sm.wbar[which]);
  };
  auto fill_ln = [&](int L, int which) {   // which = 0: self_attn_layer_norm, 1: final_layer_norm
    float* dst = which ? sm.ln2 : sm.ln1;
    fence_proxy_async();
    mbar_expect_tx(&sm.lnbar[which], 2u * HID * 4);
    bulk_g2s(dst, sm.ptab[L][2 + 2 * which], HID * 4, &sm.lnbar[which]);
    bulk_g2s(dst + HID, sm.ptab[L][3 + 2 * which], HID * 4, &sm.lnbar[which]);
  };
  auto fill_bias = [&](int buf, int L) {
    fill(sm.bias[buf], ws->bias_cta + ((size_t)L * MG_GRID + cta) * BIAS_N, BIAS_N * 2, &sm.bbar[buf]);
  };

  uint32_t par = 0;   // phase parities: bit 0..3 = weight buffers D, C, A, B; 4, 5 = ln1, ln2; 6, 7 = bias buffers
#define PAR(b) ((par >> (b)) & 1u)
#define FLIP(b) (par ^= (1u << (b)))
  int lc = 0;  // layer instances processed by this launch (alternates the bias buffers)
  if (tid == 0) {
    fill_ln(0, 0);
    fill_ln(0, 1);
    fill_bias(0, 0);
    fill_bias(1, NL > 1 ? 1 : 0);
    fill_D(0);
    fill_C(0);
    fill_A(0);
    fill_B(0);
  }

  // generation state, identical in every CTA
  int pos = a.s.pos[0], gen = a.s.gen[0], tok = a.s.tok[0], fin = a.s.finished[0];
  unsigned long long* tr = ((a.trace & 1) && cta == 0 && tid == 0) ? ws->trace : nullptr;
  int tri = 0;
#define STAMP() do { if (tr && tri < 1270) tr[tri++] = gtimer(); } while (0)
  // trace & 4: additional stamps inside the phases (tools/trace_mega.py --fine)
#define FSTAMP() do { if (tr && (a.trace & 4) && tri < 1270) tr[tri++] = gtimer(); } while (0)
  // per-CTA stamps 10..15 of (second traced step, layer NL/2): inside-phase events
  int cs_L = -1, cs_step = -1;
#define FCSTAMP(k) do { if ((a.trace & 1) && tid == 0 && cs_step == 1 && cs_L == NL / 2) ws->trace_cta[cta * 16 + (k)] = gtimer(); } while (0)
  // every CTA stamps phase k of (second traced step, layer NL/2)
  // trace & 2 (debug): every CTA records the last phase it passed, frozen once any wait has timed out
#define CSTAMP(k) do { if (tid == 0 && a.trace) { \
      if ((a.trace & 2) && !ld_volatile_i32(&ws->error)) { \
        volatile int* wh_ = ws->where + 4 * cta; wh_[0] = a.step_base + step + 1; wh_[1] = L; wh_[2] = (k); wh_[3] = lc; } \
      if ((a.trace & 1) && step == 1 && L == NL / 2) ws->trace_cta[cta * 16 + (k)] = gtimer(); } } while (0)

  float hres[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // residual stream: thread t < 256 owns elements 4t..4t+3

  // xs <- fp16(v) (threads < 256)
  auto publish_x = [&](const float* v) {
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(sm.xs + 4 * tid) = u;
  };
  // hres <- LN(hres + float(reduced vector words of this thread)), xs <- fp16(hres); all 512 threads call it
  // debug (trace & 2): lane 0 of every warp records how far it got inside residual_ln (frozen at the first time-out)
#define WMARK(m) do { if ((a.trace & 2) && lane == 0 && !ld_volatile_i32(&ws->error)) \
      *reinterpret_cast<volatile int*>(ws->wprog + cta * 16 + warp) = (int)(ep << 8) | (which << 4) | (m); } while (0)
  auto residual_ln = [&](const uint2* words, uint32_t ep, int which) {   // which = 0: ln1, 1: ln2
    const float* lnp = which ? sm.ln2 : sm.ln1;
    WMARK(1);
    if (tid < 256) {
      uint2 d;
      ll_wait_units<1>(words + 2 * tid, ep, &d, wc, ERR_Y);
      FSTAMP();
      FCSTAMP(which ? 14 : 10);
      WMARK(2);
      const __half2* hh = reinterpret_cast<const __half2*>(&d);
      const float2 p0 = __half22float2(hh[0]), p1 = __half22float2(hh[1]);
      float v[4] = {fadd(hres[0], p0.x), fadd(hres[1], p0.y), fadd(hres[2], p1.x), fadd(hres[3], p1.y)};
      mbar_wait_b(&sm.lnbar[which], PAR(4 + which), wc);
      FSTAMP();
      WMARK(3);
      layernorm_1024(v, lnp, lnp + HID, sm.red, tid);
      WMARK(4);
      hres[0] = v[0]; hres[1] = v[1]; hres[2] = v[2]; hres[3] = v[3];
      publish_x(v);
    }
    FLIP(4 + which);
    WMARK(5);
    __syncthreads();   // xs complete; every reader of the parameters is done
    WMARK(6);
  };
  // reducer CTAs: rows 8*cta..8*cta+7 of the 16 partial vectors -> balanced tree, + bias, fp16, 4 flagged words.
  // Thread t < 128: row 8*cta + (t >> 4), partial t & 15; a warp holds the row pair (2*warp, 2*warp + 1).
  // `bias` = the bias of this thread's row, read by the caller before the bias buffer may be refilled
  auto reduce_publish = [&](const uint2* pw, uint2* yw, uint32_t ep, float bias) {
    if (reducer && tid < 128) {
      const int rr = tid >> 4;
      uint32_t w;
      ll_wait_words<1>(pw + ((size_t)(8 * cta + rr)) * NG + (tid & 15), 0, ep, &w, wc, ERR_RED);
      float s = __uint_as_float(w);
      s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 1));
      s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 2));
      s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 4));
      s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 8));
      const __half y = __float2half_rn(fadd(s, bias));
      const __half yo = __shfl_xor_sync(0xffffffffu, y, 16);
      if (lane == 0) ll_store(yw + 4 * cta + warp, pack2(y, yo), ep);
    }
  };

  int step = 0;
  for (; step < a.n_steps; step++) {
    if (gen >= a.max_new || fin) break;   // uniform across the grid
    const int nkeys = pos + 1;
    const int nch = (nkeys + MA_ATTN_CHUNK - 1) / MA_ATTN_CHUNK;
    const uint32_t ep0 = (uint32_t)(a.step_base + step) * (uint32_t)NL + 1u;  // epoch of layer 0 of this step
    STAMP();

    for (int L = 0; L < NL; L++) {
      const uint32_t ep = ep0 + (uint32_t)L;
      const int bsel = lc & 1;
      cs_L = L; cs_step = step;
      const __half* lb = sm.bias[bsel];
      __half* kc = a.kv + ((size_t)(L * 2 + 0)) * NHEAD * T * HD + (size_t)g * T * HD;   // this head
      __half* vc = a.kv + ((size_t)(L * 2 + 1)) * NHEAD * T * HD + (size_t)g * T * HD;

      // ---------------- K/V prefetch into registers: first chunk of this team (rows < pos are old)
      uint4 kreg[8], vreg[8];
      const int c_first = j + GS * team;   // chunks c_first, c_first + 18, ... belong to this team
      if (c_first < nch) {
        const long base = (long)c_first * MA_ATTN_CHUNK * HD;
#pragma unroll
        for (int rho = 0; rho < 8; rho++) {
          const int r = 32 * rho + 4 * wt + grp;
          if (c_first * MA_ATTN_CHUNK + r < pos) {
            kreg[rho] = ldcg16(kc + base + (long)r * HD + 8 * li);
            vreg[rho] = ldcg16(vc + base + (long)r * HD + 8 * li);
          }
        }
      }

      // ---------------- input of the layer: token embedding (layer 0) or LN2(hres + fc2 vector) of the previous layer
      if (L == 0) {
        if (tid < 256) {
          float4 X;
          int fidx;
          if (tok < 3) {
            X = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm.gtab[2]) + (long)tok * HID + 4 * tid);
            fidx = tok;
          } else {
            const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(sm.gtab[1]) +
                                                            (long)(tok - 3) * HID + 4 * tid);
            const __half2* hh = reinterpret_cast<const __half2*>(&u);
            const float2 p0 = __half22float2(hh[0]), p1 = __half22float2(hh[1]);
            X = make_float4(p0.x, p0.y, p1.x, p1.y);
            int r = (gen - 2) % 9;
            if (r < 0) r += 9;
            fidx = r + 3;
          }
          const float4 F = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm.gtab[3]) + (long)fidx * HID + 4 * tid);
          const float4 C = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm.gtab[4]) + HID + 4 * tid);
          const float4 P = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sm.gtab[5]) + (long)(pos + 2) * HID + 4 * tid);
          hres[0] = fadd(fadd(fadd(X.x, F.x), C.x), P.x);
          hres[1] = fadd(fadd(fadd(X.y, F.y), C.y), P.y);
          hres[2] = fadd(fadd(fadd(X.z, F.z), C.z), P.z);
          hres[3] = fadd(fadd(fadd(X.w, F.w), C.w), P.w);
          publish_x(hres);
        }
        __syncthreads();
      } else {
        residual_ln(ws->yb_w, ep - 1, 1);   // the fc2 vector of layer L-1
        if (tid == 0) fill_ln(L, 1);
      }
      STAMP();
      CSTAMP(0);

      // ---------------- qkv phase: warps 0..n_qp-1 own one row pair each
      mbar_wait_b(&sm.bbar[bsel], PAR(6 + bsel), wc);
      FLIP(6 + bsel);
      // Only the warps that read a weight buffer wait for it: a warp that waits without being counted by
      // last_warp_out could still be waiting for phase p when the refill (phase p + 1) completes, and would then
      // wait for phase p + 2 forever (parity aliasing).  Every thread flips its parity bit once per phase.
      if (warp < 11) mbar_wait_b(&sm.wbar[0], PAR(0), wc);
      FLIP(0);
      FSTAMP();
      if (warp < sm.n_qp) {
        const __half* w2[2] = {bufD + (size_t)(2 * warp) * HID, bufD + (size_t)(2 * warp + 1) * HID};
        float acc[2];
        gemv_k1024<2>(w2, sm.xs, lane, acc);
        acc[0] = warp_sum(acc[0]);
        acc[1] = warp_sum(acc[1]);
        if (lane == 0) {
          const __half h0 = __float2half_rn(fadd(acc[0], __half2float(lb[BIAS_QKV + 2 * warp])));
          const __half h1 = __float2half_rn(fadd(acc[1], __half2float(lb[BIAS_QKV + 2 * warp + 1])));
          const uint32_t word = pack2(h0, h1);
          const int p = sm.qp0 + warp;
          ll_store(ws->qkv_w + g * 96 + p, word, ep);
          if (sm.qsec > 0)   // k / v of the current token also go to the cache for later steps
            *reinterpret_cast<uint32_t*>((sm.qsec == 1 ? kc : vc) + (long)pos * HD + 2 * (p & 31)) = word;
        }
      }
      if (warp < 11 && last_warp_out(&sm.cnt[0], 11, lane)) {
        if (L + 1 < NL) fill_D(L + 1); else fill_lm(0);
      }
      // Warps without rows wait HERE (blocked at the barrier) rather than run ahead into the next phase's polls: a
      // spinning warp shares its scheduler and the SM's load/store pipeline with the warps that still compute, and
      // was measured to stretch their phase by up to 3x (profiles/mega_trace_r02.txt)
      __syncthreads();
      STAMP();
      CSTAMP(1);

      // ---------------- attention phase: chunk c of head g belongs to team (CTA c % 9, team (c / 9) & 1) of the group
      {
        // q of this head: flagged words of the group's exchange buffer (once per layer, before the chunk loop)
        uint4 qp = make_uint4(0u, 0u, 0u, 0u);
        if (c_first < nch) {
          uint2 d[2];
          ll_wait_units<2>(ws->qkv_w + g * 96 + 4 * li, ep, d, wc, ERR_QKV);
          qp = make_uint4(d[0].x, d[0].y, d[1].x, d[1].y);
        }
        int it = 0;
        for (int c = c_first; c < nch; c += 2 * GS, it++) {
          const int len = min(MA_ATTN_CHUNK, nkeys - c * MA_ATTN_CHUNK);
          const long base = (long)c * MA_ATTN_CHUNK * HD;
          const int cur = pos - c * MA_ATTN_CHUNK;  // row of the current token inside this chunk (if 0 <= cur < 256)
#pragma unroll
          for (int rho = 0; rho < 8; rho++) {
            const int r = 32 * rho + 4 * wt + grp;
            if (r < len && it > 0 && r != cur) {  // later chunks were not prefetched into registers (L2 prefetch only)
              kreg[rho] = ldcg16(kc + base + (long)r * HD + 8 * li);
              vreg[rho] = ldcg16(vc + base + (long)r * HD + 8 * li);
            }
          }
          // for the chunk that holds it, k / v of the current token come from the exchange buffer too
          if (cur >= 0 && cur < MA_ATTN_CHUNK) {
            const int rho_c = cur >> 5, gl_c = cur & 31;
            if (4 * wt + grp == gl_c) {
              uint2 kk[2], vv[2];
              ll_wait_units<2>(ws->qkv_w + g * 96 + 32 + 4 * li, ep, kk, wc, ERR_QKV);
              ll_wait_units<2>(ws->qkv_w + g * 96 + 64 + 4 * li, ep, vv, wc, ERR_QKV);
#pragma unroll
              for (int rho = 0; rho < 8; rho++)
                if (rho == rho_c) {
                  kreg[rho] = make_uint4(kk[0].x, kk[0].y, kk[1].x, kk[1].y);
                  vreg[rho] = make_uint4(vv[0].x, vv[0].y, vv[1].x, vv[1].y);
                }
            }
          }
          float sreg[8];
          float lmax = -INFINITY;
#pragma unroll
          for (int rho = 0; rho < 8; rho++) {
            const int r = 32 * rho + 4 * wt + grp;
            float p = dot8(qp, kreg[rho], 0.0f);
            p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 4));
            p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 2));
            p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 1));
            sreg[rho] = fmul(p, 0.125f);
            if (r < len) lmax = fmaxf(lmax, sreg[rho]);
          }
          lmax = warp_max(lmax);
          team_sync(team);  // previous users of wmax / ared of this team are done
          if (lane == 0) sm.wmax[team][wt] = lmax;
          team_sync(team);
          float cmax = sm.wmax[team][0];
#pragma unroll
          for (int w2 = 1; w2 < 8; w2++) cmax = fmaxf(cmax, sm.wmax[team][w2]);
          float l = 0.0f, o[8];
#pragma unroll
          for (int jj = 0; jj < 8; jj++) o[jj] = 0.0f;
#pragma unroll
          for (int rho = 0; rho < 8; rho++) {
            const int r = 32 * rho + 4 * wt + grp;
            if (r < len) {
              const float e = ma_exp(fsub(sreg[rho], cmax));
              l = fadd(l, e);
              pv8(__float2half_rn(e), vreg[rho], o);
            }
          }
          l = fadd(l, __shfl_xor_sync(0xffffffffu, l, 16));
          l = fadd(l, __shfl_xor_sync(0xffffffffu, l, 8));
#pragma unroll
          for (int jj = 0; jj < 8; jj++) {
            o[jj] = fadd(o[jj], __shfl_xor_sync(0xffffffffu, o[jj], 16));
            o[jj] = fadd(o[jj], __shfl_xor_sync(0xffffffffu, o[jj], 8));
          }
          if (grp == 0) {
#pragma unroll
            for (int jj = 0; jj < 8; jj++) sm.ared[team][wt][8 * li + jj] = o[jj];
            if (li == 0) sm.ared[team][wt][64] = l;
          }
          team_sync(team);
          if (tl < 65) {
            float x[8];
#pragma unroll
            for (int w2 = 0; w2 < 8; w2++) x[w2] = sm.ared[team][w2][tl];
            const float rsum = fadd(fadd(fadd(x[0], x[1]), fadd(x[2], x[3])), fadd(fadd(x[4], x[5]), fadd(x[6], x[7])));
            uint2* part = ws->part_w + ((long)g * MAX_CHUNKS + c) * PARTF;
            ll_store(part + (tl < 64 ? tl : 65), __float_as_uint(rsum), ep);
            if (tl == 64) ll_store(part + 64, __float_as_uint(cmax), ep);
          }
        }
        // the same chunks of the next layer instance: start pulling them into L2 now
        if (tl == 0) {
          const int Ln = (L + 1 < NL) ? L + 1 : 0;
          const __half* kn = a.kv + ((size_t)(Ln * 2 + 0)) * NHEAD * T * HD + (size_t)g * T * HD;
          const __half* vn = a.kv + ((size_t)(Ln * 2 + 1)) * NHEAD * T * HD + (size_t)g * T * HD;
          for (int c = c_first; c < nch; c += 2 * GS) {
            const uint32_t bytes = (uint32_t)min(MA_ATTN_CHUNK, nkeys - c * MA_ATTN_CHUNK) * HD * 2;
            l2_prefetch(kn + (long)c * MA_ATTN_CHUNK * HD, bytes);
            l2_prefetch(vn + (long)c * MA_ATTN_CHUNK * HD, bytes);
          }
        }
      }
      __syncthreads();   // a team without chunks waits here, not in the polls below, while the other team computes
      STAMP();
      CSTAMP(2);

      // ---------------- merge of the chunk partials of this head, redundantly in every CTA of the group: all 66 words
      // of every chunk are staged in shared memory by all threads with their polls in flight together; the chunk
      // weights exp(m_c - M) are computed once per chunk; 64 threads run the ascending fma chain of the canonical merge
#define WMARK2(m) do { if ((a.trace & 2) && lane == 0 && !ld_volatile_i32(&ws->error)) \
      *reinterpret_cast<volatile int*>(ws->wprog + cta * 16 + warp) = (int)(ep << 8) | (m); } while (0)
      WMARK2(7);
      {
        const uint2* part = ws->part_w + (long)g * MAX_CHUNKS * PARTF;
        const int nw = nch * PARTF;
        for (int i0 = 0; i0 < nw; i0 += 4 * MG_THREADS) {
          uint2 w[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int i = i0 + tid + MG_THREADS * k;
            if (i < nw) w[k] = ll_load1(part + i);
          }
          unsigned n = 0;
          unsigned long long t0 = 0;
          for (;;) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int i = i0 + tid + MG_THREADS * k;
              if (i < nw && w[k].y != ep) {
                ok = false;
                w[k] = ll_load1(part + i);
              }
            }
            if (ok) break;
            if (wc.backoff_ns) __nanosleep(wc.backoff_ns);
            if ((++n & 255u) == 0 && wait_slow(wc, t0, ERR_PART)) break;
          }
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int i = i0 + tid + MG_THREADS * k;
            if (i < nw) sm.part[i] = __uint_as_float(w[k].x);
          }
        }
      }
      FSTAMP();
      FCSTAMP(15);
      WMARK2(8);
      __syncthreads();
      WMARK2(9);
      if (tid < 64) {
        float M = -INFINITY;
        for (int cc = 0; cc < nch; cc++) M = fmaxf(M, sm.part[cc * PARTF + 64]);
        for (int cc = tid; cc < nch; cc += 64) sm.mw[cc] = ma_exp(fsub(sm.part[cc * PARTF + 64], M));
        merge_sync();
        WMARK2(10);
        float Lsum = 0.0f, O = 0.0f;
        for (int cc = 0; cc < nch; cc++) {
          const float wgt = sm.mw[cc];
          Lsum = ffma(sm.part[cc * PARTF + 65], wgt, Lsum);
          O = ffma(sm.part[cc * PARTF + tid], wgt, O);
        }
        sm.ab[tid] = __float2half_rn(__fdiv_rn(O, Lsum));
      }
      WMARK2(11);
      __syncthreads();
      WMARK2(12);
      STAMP();
      CSTAMP(3);

      // ---------------- out_proj, split-K by head: this CTA's rows x the 64 columns of head g; one row per 8 lanes
      mbar_wait_b(&sm.wbar[1], PAR(1), wc);
      FLIP(1);
      FSTAMP();
      {
        const uint4 av = *reinterpret_cast<const uint4*>(sm.ab + 8 * li);
        const bool withhold = a.fault && cta == a.fault - 1 && a.step_base + step >= 2;
#pragma unroll
        for (int it = 0; it < 2; it++) {
          const int r = 64 * it + 4 * warp + grp;
          const int rc = min(r, sm.n_sr - 1);
          float p = dot8(*reinterpret_cast<const uint4*>(bufC + rc * HD + 8 * li), av, 0.0f);
          p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 4));
          p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 2));
          p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 1));
          if (li == 0 && r < sm.n_sr && !withhold) ll_store(ws->pa_w + (size_t)(sm.sr0 + r) * NG + g, __float_as_uint(p), ep);
        }
        if (last_warp_out(&sm.cnt[1], MG_WARPS, lane)) {
          if (L + 1 < NL) fill_C(L + 1); else fill_lm(1);
        }
      }
      STAMP();
      CSTAMP(4);
      reduce_publish(ws->pa_w, ws->ya_w, ep, __half2float(lb[BIAS_OUT + ((tid >> 4) & 7)]));
      STAMP();
      CSTAMP(5);

      // ---------------- fc1 phase: input = LN1(hres + out_proj vector); warps 0..n_fp-1 own one row pair each
      residual_ln(ws->ya_w, ep, 0);
      if (tid == 0) fill_ln((L + 1) % NL, 0);
      STAMP();
      CSTAMP(6);
      if (warp < 15) mbar_wait_b(&sm.wbar[2], PAR(2), wc);
      FLIP(2);
      FSTAMP();
      FCSTAMP(11);
      if (warp < sm.n_fp) {
        const __half* w2[2] = {bufA + (size_t)(2 * warp) * HID, bufA + (size_t)(2 * warp + 1) * HID};
        float acc[2];
        gemv_k1024<2>(w2, sm.xs, lane, acc);
        acc[0] = warp_sum(acc[0]);
        acc[1] = warp_sum(acc[1]);
        if (lane == 0) {
          __half h0 = __float2half_rn(fadd(acc[0], __half2float(lb[BIAS_FC1 + 2 * warp])));
          __half h1 = __float2half_rn(fadd(acc[1], __half2float(lb[BIAS_FC1 + 2 * warp + 1])));
          if (__half2float(h0) < 0.0f) h0 = __float2half_rn(0.0f);
          if (__half2float(h1) < 0.0f) h1 = __float2half_rn(0.0f);
          ll_store(ws->f_w + g * 128 + sm.fp0 + warp, pack2(h0, h1), ep);
        }
      }
      if (warp < 15 && last_warp_out(&sm.cnt[2], 15, lane)) {
        if (L + 1 < NL) fill_A(L + 1); else fill_lm(2);
      }
      __syncthreads();   // as above: nobody polls for the fc1 activations while a warp of this CTA still computes its rows
      STAMP();
      CSTAMP(7);

      // ---------------- fc2, split-K by group: this CTA's rows x the 256 fc1 activations of group g; 8 rows per warp
      const float bias_fc2 = __half2float(lb[BIAS_FC2 + ((tid >> 4) & 7)]);   // (reducers) before the buffer is refilled
      {
        uint4 xr;
        {
          uint2 d[2];
          ll_wait_units<2>(ws->f_w + g * 128 + 4 * lane, ep, d, wc, ERR_F);
          xr = make_uint4(d[0].x, d[0].y, d[1].x, d[1].y);
        }
        FSTAMP();
        FCSTAMP(12);
        mbar_wait_b(&sm.wbar[3], PAR(3), wc);
        FLIP(3);
        FSTAMP();
        FCSTAMP(13);
        if (8 * warp < sm.n_sr) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int rc = min(8 * warp + i, sm.n_sr - 1);
            v[i] = dot8(*reinterpret_cast<const uint4*>(bufB + (size_t)rc * 256 + 8 * lane), xr, 0.0f);
          }
          // transposing butterfly over lane bits 16, 8, 4 (row bit 2, 1, 0), then plain xor-2, xor-1: the canonical
          // xor-16,8,4,2,1 sum of row ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)
#pragma unroll
          for (int s = 4; s >= 1; s >>= 1) {
            const bool up = (lane & (4 * s)) != 0;
#pragma unroll
            for (int i = 0; i < s; i++) {
              const float mine = up ? v[i + s] : v[i];
              const float other = up ? v[i] : v[i + s];
              v[i] = fadd(mine, __shfl_xor_sync(0xffffffffu, other, 4 * s));
            }
          }
          float p = v[0];
          p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 2));
          p = fadd(p, __shfl_xor_sync(0xffffffffu, p, 1));
          const int r = 8 * warp + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
          if ((lane & 3) == 0 && r < sm.n_sr) ll_store(ws->pb_w + (size_t)(sm.sr0 + r) * NG + g, __float_as_uint(p), ep);
        }
        if (last_warp_out(&sm.cnt[3], MG_WARPS, lane)) {
          fill_B((L + 1 < NL) ? L + 1 : 0);
          // the bias buffer of this layer instance is free: refill it for the instance that uses it next (L + 2)
          fill_bias(bsel, (L + 2) % NL);
        }
      }
      STAMP();
      CSTAMP(8);
      reduce_publish(ws->pb_w, ws->yb_w, ep, bias_fc2);
      lc++;
      STAMP();
      CSTAMP(9);
    }

    // ---------------- lm_head on LN2 of the last layer + greedy pick
    const uint32_t epc = (uint32_t)(a.step_base + step) + 1u;
    residual_ln(ws->yb_w, ep0 + (uint32_t)NL - 1u, 1);
    if (tid == 0) fill_ln(0, 1);
    mbar_wait_b(&sm.wbar[0], PAR(0), wc);
    mbar_wait_b(&sm.wbar[1], PAR(1), wc);
    mbar_wait_b(&sm.wbar[2], PAR(2), wc);
    FLIP(0); FLIP(1); FLIP(2);
    {
      // rows warp, warp + 16, warp + 32, warp + 48 together
      const __half* w4[4];
      bool has[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int r = warp + MG_WARPS * i;
        has[i] = r < sm.n_lm;
        w4[i] = bufD + (size_t)(has[i] ? r : 0) * HID;
      }
      float acc[4];
      gemv_k1024<4>(w4, sm.xs, lane, acc);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        acc[i] = warp_sum(acc[i]);
        if (lane == 0 && has[i]) sm.stage16[warp + MG_WARPS * i] = __float2half_rn(acc[i]);
      }
    }
    __syncthreads();
    float bestv = -INFINITY;
    int besti = 0x7fffffff;
    if (tid < sm.n_lm) {
      const __half hv = sm.stage16[tid];
      if (a.logits_out) a.logits_out[(long)gen * vocab + sm.row0_lm + tid] = hv;
      bestv = __half2float(hv);
      besti = sm.row0_lm + tid;
    }
    if (warp < 2) {  // sm.n_lm <= 58: the candidates live in the first two warps
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
      }
      if (lane == 0) { sm.bval[warp] = bestv; sm.bidx[warp] = besti; }
    }
    __syncthreads();
    if (tid == 0) {
      float bv = sm.bval[0];
      int bi = sm.bidx[0];
      if (sm.bval[1] > bv || (sm.bval[1] == bv && sm.bidx[1] < bi)) { bv = sm.bval[1]; bi = sm.bidx[1]; }
      __threadfence();  // publish this step's KV-cache rows before the step's final hand-off
      ll_store(ws->cand_w + 2 * cta, __float_as_uint(bv), epc);
      ll_store(ws->cand_w + 2 * cta + 1, (uint32_t)bi, epc);
      // weights of the next token's first layer (B was refilled after the last fc2)
      fill_D(0);
      fill_C(0);
      fill_A(0);
    }
    {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      if (tid < MG_GRID) {
        uint2 d;
        ll_wait_units<1>(ws->cand_w + 2 * tid, epc, &d, wc, ERR_CAND);
        bv = __uint_as_float(d.x);
        bi = (int)d.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      __syncthreads();
      if (lane == 0) { sm.bval[warp] = bv; sm.bidx[warp] = bi; }
      if (tid == 0) sm.errflag = ld_volatile_i32(&ws->error);   // one reader: the decision below is CTA-uniform
      __syncthreads();
      bv = sm.bval[0];
      bi = sm.bidx[0];
      for (int w2 = 1; w2 < MG_WARPS; w2++)
        if (sm.bval[w2] > bv || (sm.bval[w2] == bv && sm.bidx[w2] < bi)) { bv = sm.bval[w2]; bi = sm.bidx[w2]; }
      if (sm.errflag) break;   // a wait of this step failed somewhere: emit nothing more
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

  // every weight / parameter buffer has a refill in flight here: drain them before the shared memory is released
  // (after a failed wait some of them may never have been issued: the waits are bounded)
  for (int b = 0; b < 4; b++) mbar_wait_b(&sm.wbar[b], PAR(b), wc);
  for (int b = 0; b < 2; b++) mbar_wait_b(&sm.lnbar[b], PAR(4 + b), wc);
  for (int b = 0; b < 2; b++) mbar_wait_b(&sm.bbar[b], PAR(6 + b), wc);
  if (cta == 0 && tid == 0) {
    const int e = ld_volatile_i32(&ws->error);
    a.s.pos[0] = pos; a.s.gen[0] = gen; a.s.tok[0] = tok; a.s.finished[0] = e ? 1 : fin;
    if (e) a.s.lens[0] = -1;   // surfaced by the callers of ma_decode_generate (out_lens): no silently wrong mesh
    if (a.nkeys_next) *a.nkeys_next = pos + 1;
    if (a.all_done) *a.all_done = e ? 1 : fin;
  }
}

// ---- host side -----------------------------------------------------------------------------------
size_t mega_workspace_bytes() { return sizeof(MegaWs) + 256; }

// bias_cta[L][cta][BIAS_N] <- the biases of the rows CTA `cta` = (group g, rank j) owns in layer L
__global__ void mega_pack_bias_kernel(MegaWs* ws) {
  const int L = blockIdx.y, cta = blockIdx.x, t = threadIdx.x;  // BIAS_N threads
  const int g = cta / GS, j = cta % GS;
  const ma_decoder_weights& W = ws->w;
  __half v = __float2half_rn(0.0f);
  if (t < BIAS_FC1) {
    const int qp0 = qkv_pair0(j), n = 2 * (qkv_pair0(j + 1) - qp0);
    if (t < n) v = reinterpret_cast<const __half*>(W.bqkv[L])[(qp0 >> 5) * HID + g * HD + 2 * (qp0 & 31) + t];
  } else if (t < BIAS_OUT) {
    const int i = t - BIAS_FC1, fp0 = fc1_pair0(j), n = 2 * (fc1_pair0(j + 1) - fp0);
    if (i < n) v = reinterpret_cast<const __half*>(W.b1[L])[g * 256 + 2 * fp0 + i];
  } else if (t < BIAS_FC2) {
    if (cta < NRED) v = reinterpret_cast<const __half*>(W.bo[L])[8 * cta + (t - BIAS_OUT)];
  } else if (t < BIAS_FC2 + 8) {
    if (cta < NRED) v = reinterpret_cast<const __half*>(W.b2[L])[8 * cta + (t - BIAS_FC2)];
  }
  ws->bias_cta[((size_t)L * MG_GRID + cta) * BIAS_N + t] = v;
}

// wo_p[L][head][row][0..63] <- W_o[row][64 head ..];  w2_p[L][group][row][0..255] <- W_2[row][256 group ..]
__global__ void mega_pack_weights_kernel(MegaWs* ws) {
  const int L = blockIdx.y, row = blockIdx.x, t = threadIdx.x;  // 128 threads, one block per output row
  const ma_decoder_weights& W = ws->w;
  {  // W_o row: 128 uint4; piece i = columns 8i.. -> head i / 8, offset i % 8
    const uint4* src = reinterpret_cast<const uint4*>(W.wo[L]) + (size_t)row * (HID / 8);
    uint4* dst = reinterpret_cast<uint4*>(ws->wo_p + (size_t)L * HID * HID);
    const int hd = t >> 3, q = t & 7;
    dst[((size_t)hd * HID + row) * (HD / 8) + q] = src[t];
  }
  {  // W_2 row: 512 uint4; piece i -> group i / 32, offset i % 32
    const uint4* src = reinterpret_cast<const uint4*>(W.w2[L]) + (size_t)row * (FFN / 8);
    uint4* dst = reinterpret_cast<uint4*>(ws->w2_p + (size_t)L * HID * FFN);
    for (int i = t; i < FFN / 8; i += 128) {
      const int gg = i >> 5, q = i & 31;
      dst[((size_t)gg * HID + row) * 32 + q] = src[i];
    }
  }
}

static int g_mega_ok = -1;   // 1: 144 CTAs of this shape are co-resident on this device
static size_t mega_smem_bytes() { return (size_t)BYTES_D + BYTES_C + BYTES_A + BYTES_B + sizeof(MegaSmem); }

// Can the persistent kernel run here?  (shared memory per CTA, one CTA on each of >= 144 SMs)
int mega_supported() {
  if (g_mega_ok >= 0) return g_mega_ok;
  g_mega_ok = 0;
  const size_t smem = mega_smem_bytes();
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("mega: cannot get %zu bytes of shared memory", smem);
    cudaGetLastError();
    return 0;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_mega_kernel, MG_THREADS, smem) != cudaSuccess ||
      per_sm * sms < MG_GRID) {
    set_error("mega: %d SMs x %d resident CTAs < %d", sms, per_sm, MG_GRID);
    cudaGetLastError();
    return 0;
  }
  g_mega_ok = 1;
  return 1;
}

int mega_prepare(const ma_decoder_weights* w, void* mega_ws, cudaStream_t st) {
  MegaWs* ws = reinterpret_cast<MegaWs*>(mega_ws);
  if (w->n_layers > MA_MAX_LAYERS || w->vocab > ROWS_LM * MG_GRID) {
    set_error("mega: n_layers=%d / vocab=%d out of range", w->n_layers, w->vocab);
    return 1;
  }
  if (cudaMemsetAsync(ws, 0, offsetof(MegaWs, w), st) != cudaSuccess) return 1;  // all epochs 0, error 0
  if (cudaMemcpyAsync(&ws->w, w, sizeof(ma_decoder_weights), cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
  mega_pack_bias_kernel<<<dim3(MG_GRID, w->n_layers), BIAS_N, 0, st>>>(ws);
  mega_pack_weights_kernel<<<dim3(HID, w->n_layers), 128, 0, st>>>(ws);
  count_launch(2);
  return check_launch("mega_pack_kernels") ? 0 : 1;
}

int mega_error_flag_offset() { return (int)offsetof(MegaWs, error); }
int mega_trace_offset() { return (int)offsetof(MegaWs, trace); }
int mega_trace_cta_offset() { return (int)offsetof(MegaWs, trace_cta); }
int mega_where_offset() { return (int)offsetof(MegaWs, where); }
int mega_fail_offset() { return (int)offsetof(MegaWs, fail); }
int mega_wprog_offset() { return (int)offsetof(MegaWs, wprog); }
int mega_ws_offset(int which) {   // offsets of the exchange buffers (decoding MegaWs::fail[5])
  switch (which) {
    case 0: return (int)offsetof(MegaWs, qkv_w);
    case 1: return (int)offsetof(MegaWs, f_w);
    case 2: return (int)offsetof(MegaWs, pa_w);
    case 3: return (int)offsetof(MegaWs, pb_w);
    case 4: return (int)offsetof(MegaWs, ya_w);
    case 5: return (int)offsetof(MegaWs, yb_w);
    case 6: return (int)offsetof(MegaWs, cand_w);
    case 7: return (int)offsetof(MegaWs, part_w);
    default: return (int)offsetof(MegaWs, error);
  }
}

static unsigned long long g_mega_timeout_ns = 2000000000ull;   // 2 s per wait
static int g_mega_fault = 0;
void mega_set_debug(unsigned long long timeout_ns, int fault) {
  if (timeout_ns) g_mega_timeout_ns = timeout_ns;
  g_mega_fault = fault;
}

bool mega_fits(int tmax) { return tmax <= MAX_CHUNKS * MA_ATTN_CHUNK; }

int mega_enqueue(const ma_decoder_weights* w, SeqState s, int tmax, __half* kv, void* mega_ws, const SampleArgs& sa,
                 int n_steps, int step_base, int trace, cudaStream_t st) {
  if (!mega_fits(tmax)) {
    set_error("mega: tmax=%d exceeds %d keys", tmax, MAX_CHUNKS * MA_ATTN_CHUNK);
    return 1;
  }
  if (!mega_supported()) return 1;
  (void)w;
  MegaArgs a;
  memset(&a, 0, sizeof(a));
  a.ws = reinterpret_cast<MegaWs*>(mega_ws);
  a.s = s;
  a.kv = kv;
  a.T = tmax;
  a.n_steps = n_steps;
  a.step_base = step_base;
  a.max_new = sa.max_new; a.eos_id = sa.eos_id; a.pad_id = sa.pad_id;
  a.out_ids = sa.out_ids; a.forced = sa.forced; a.logits_out = sa.logits_out; a.all_done = sa.all_done;
  a.nkeys_next = sa.nkeys_next;
  a.trace = trace;
  a.fault = g_mega_fault;
  a.timeout_ns = g_mega_timeout_ns;
  {
    static int bo = -1;
    if (bo < 0) { const char* e = getenv("MA_MEGA_BACKOFF_NS"); bo = e ? atoi(e) : 0; }
    a.backoff_ns = (unsigned)bo;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(MG_GRID);
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = mega_smem_bytes();
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
