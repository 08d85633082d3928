// attention_tc.cu -- dense (non-causal) attention softmax(q K^T * scale) V on the 5th-generation tensor cores, for the
// stages compared under a tolerance: the Perceiver cross-attention (257 queries x 4096 keys x 12 heads,
// transformer_blocks.py:166-185), the 24 Michelangelo self-attention layers (:57-74) and the 6 BERT layers of the
// detokenizer (meshanything.py:50-80).  The decoder keeps the canonical CUDA-core attention.
//
// One CTA = 128 queries of one (slot, head); loop over blocks of 128 keys (flash-attention recurrence):
//   S = Q K_j^T          tcgen05.mma M128 N128 K16 x4   (Q, K_j: K-major tiles [128][64 halfs], TMA, 128-byte swizzle)
//   P = exp2((S - m) * scale * log2 e)  one thread per query row: tcgen05.ld of its S row from TMEM, running max /
//                        sum, P rounded to fp16 and stored to shared memory in the same swizzled K-major layout
//   O_j = P V_j          tcgen05.mma M128 N64 K16 x8    (P: two [128][64] tiles; V_j^T: two [64 d][64 keys] tiles --
//                        V is kept TRANSPOSED in global memory ([slot][head][64][Tpad]) so that it is K-major too)
//   O = O * alpha + O_j  in registers (64 fp32 per thread), so nothing in TMEM is ever rescaled
// 192 threads: warps 0-3 softmax/epilogue (TMEM lane quadrant = warp), warp 4 MMA issuer + TMEM allocator,
// warp 5 TMA producer.  K/V^T double-buffered; S (128 columns) and O_j (64 columns) live in one 256-column TMEM
// allocation.  mbarriers: q_full, kv_full/kv_empty[2], s_full, p_full (128 arrivals), o_full, o_empty (128).
#include "internal.h"
#include "tc_common.cuh"

namespace ma {

constexpr int FA_BQ = 128, FA_BK = 128, FA_STAGES = 2, FA_THREADS = 192;
constexpr uint32_t FA_TMEM_COLS = 256, FA_S_COL = 0, FA_O_COL = 128;
constexpr uint32_t FA_SPIN_LIMIT = 1u << 27;  // bounded polls (a few seconds): a protocol bug traps instead of hanging the GPU

struct alignas(1024) FaSmem {
  __half q[FA_BQ * 64];                   // 16 KB
  __half k[FA_STAGES][FA_BK * 64];        // 16 KB each
  __half vt[FA_STAGES][2][64 * 64];       // per stage: V^T for keys 0..63 and 64..127 of the block, 8 KB each
  __half p[2][FA_BQ * 64];                // P tile, keys 0..63 | 64..127
  uint64_t q_full, kv_full[FA_STAGES], kv_empty[FA_STAGES], s_full, p_full, o_full, o_empty;
  uint32_t tmem_base;
};

__device__ __forceinline__ void fa_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < FA_SPIN_LIMIT; spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

struct FaArgs {
  __half* out;
  int ldo, H, rows_per_slot, nkeys;
  long T;       // key capacity per (slot, head) in the K tensor
  float sl2;    // scale * log2(e)
};

__global__ void __launch_bounds__(FA_THREADS, 1)
    attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_vt, FaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FaSmem& sm = *reinterpret_cast<FaSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, slot = blockIdx.z;
  const int row0 = slot * a.rows_per_slot + qt * FA_BQ;          // first query row of this tile (global row index)
  const int valid_rows = min(FA_BQ, a.rows_per_slot - qt * FA_BQ);
  const int nb = (a.nkeys + FA_BK - 1) / FA_BK;
  const long head = (long)slot * a.H + h;

  if (warp == 5 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_vt) : "memory");
    mbar_init(&sm.q_full, 1);
    for (int s = 0; s < FA_STAGES; s++) {
      mbar_init(&sm.kv_full[s], 1);
      mbar_init(&sm.kv_empty[s], 1);
    }
    mbar_init(&sm.s_full, 1);
    mbar_init(&sm.p_full, FA_BQ);
    mbar_init(&sm.o_full, 1);
    mbar_init(&sm.o_empty, FA_BQ);
    mbar_fence_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                 "n"(FA_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 5) {
    // ---------------- TMA producer
    if (elect_one()) {
      mbar_expect_tx(&sm.q_full, FA_BQ * 64 * 2);
      tma_load_2d(sm.q, &map_q, 64 * h, row0, &sm.q_full);
      for (int j = 0; j < nb; j++) {
        const int s = j % FA_STAGES;
        const uint32_t ph = (j / FA_STAGES) & 1;
        fa_wait(&sm.kv_empty[s], ph ^ 1);
        mbar_expect_tx(&sm.kv_full[s], (FA_BK * 64 + 2 * 64 * 64) * 2);
        tma_load_2d(sm.k[s], &map_k, 0, (int)(head * a.T + (long)j * FA_BK), &sm.kv_full[s]);
        tma_load_2d(sm.vt[s][0], &map_vt, j * FA_BK, (int)(head * 64), &sm.kv_full[s]);
        tma_load_2d(sm.vt[s][1], &map_vt, j * FA_BK + 64, (int)(head * 64), &sm.kv_full[s]);
      }
    }
  } else if (warp == 4) {
    // ---------------- MMA issuer.  idesc: D=f32, A=B=f16, both K-major, N>>3 at bit 17, M>>4 at bit 24
    constexpr uint32_t idesc_qk = (1u << 4) | ((uint32_t)(FA_BK >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
    constexpr uint32_t idesc_pv = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
    auto issue_qk = [&](int s) {
      const uint64_t ad = umma_desc(sm.q), bd = umma_desc(sm.k[s]);
#pragma unroll
      for (int k = 0; k < 4; k++) umma_f16(tmem + FA_S_COL, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_qk, k ? 1u : 0u);
      umma_commit(&sm.s_full);
    };
    fa_wait(&sm.q_full, 0);
    fa_wait(&sm.kv_full[0], 0);
    tc_fence_after();
    if (elect_one()) issue_qk(0);
    __syncwarp();
    for (int j = 0; j < nb; j++) {
      const int s = j % FA_STAGES;
      fa_wait(&sm.p_full, j & 1);                    // P_j is in shared memory, S_j has been consumed
      if (j > 0) fa_wait(&sm.o_empty, (j - 1) & 1);  // O_{j-1} has been read out of TMEM
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          const uint64_t ad = umma_desc(sm.p[kk >> 2]) + (uint64_t)((kk & 3) * 2);
          const uint64_t bd = umma_desc(sm.vt[s][kk >> 2]) + (uint64_t)((kk & 3) * 2);
          umma_f16(tmem + FA_O_COL, ad, bd, idesc_pv, kk ? 1u : 0u);
        }
        umma_commit(&sm.o_full);       // O_j complete (also: P and this K/V stage are free)
        umma_commit(&sm.kv_empty[s]);
      }
      __syncwarp();
      if (j + 1 < nb) {
        const int s2 = (j + 1) % FA_STAGES;
        fa_wait(&sm.kv_full[s2], ((j + 1) / FA_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) issue_qk(s2);
        __syncwarp();
      }
    }
  } else {
    // ---------------- softmax + output: thread = query row r of the tile = TMEM lane r
    const int r = 32 * warp + lane;
    const uint32_t trow = tmem + ((uint32_t)(32 * warp) << 16);
    float m_run = -INFINITY, l_run = 0.0f;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; i++) o[i] = 0.0f;
    unsigned char* const prow0 = reinterpret_cast<unsigned char*>(sm.p[0]) + r * 128;
    unsigned char* const prow1 = reinterpret_cast<unsigned char*>(sm.p[1]) + r * 128;
    const int sw = r & 7;  // 128-byte swizzle: 16-byte chunk c of row r lives at chunk c ^ (r % 8)
    for (int j = 0; j < nb; j++) {
      const int nvalid = min(FA_BK, a.nkeys - j * FA_BK);
      fa_wait(&sm.s_full, j & 1);
      tc_fence_after();
      uint32_t v[32];
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; c++) {
        tmem_ld32(trow + FA_S_COL + 32 * c, v);
#pragma unroll
        for (int i = 0; i < 32; i++)
          if (32 * c + i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx);   // nvalid >= 1, so m_new is finite
      const float alpha = (m_run == -INFINITY) ? 0.0f : exp2f((m_run - m_new) * a.sl2);
      float psum = 0.0f;
      // (PV_{j-1} finished reading the P tile before this thread folded O_{j-1} in -- o_full below -- so it is free)
#pragma unroll 1
      for (int c = 0; c < 4; c++) {
        tmem_ld32(trow + FA_S_COL + 32 * c, v);
        __half ph[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float p = (32 * c + i < nvalid) ? exp2f((__uint_as_float(v[i]) - m_new) * a.sl2) : 0.0f;
          ph[i] = __float2half_rn(p);
          psum += __half2float(ph[i]);
        }
        // columns 32c .. 32c+31 = 16-byte chunks 4(c%2) .. 4(c%2)+3 of row r in tile c/2
        unsigned char* base = (c < 2) ? prow0 : prow1;
#pragma unroll
        for (int g = 0; g < 4; g++) {
          const int chunk = (4 * (c & 1) + g) ^ sw;
          *reinterpret_cast<uint4*>(base + 16 * chunk) = reinterpret_cast<const uint4*>(ph)[g];
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&sm.p_full);
      l_run = l_run * alpha + psum;
      m_run = m_new;
      // O = O * alpha + O_j   (O_{j-1} was folded in during the previous iteration, right here)
      fa_wait(&sm.o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; c++) {
        tmem_ld32(trow + FA_O_COL + 32 * c, v);
#pragma unroll
        for (int i = 0; i < 32; i++) o[32 * c + i] = fmaf(o[32 * c + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
      mbar_arrive(&sm.o_empty);
    }
    if (r < valid_rows) {
      const float inv = 1.0f / l_run;
      __half out[64];
#pragma unroll
      for (int i = 0; i < 64; i++) out[i] = __float2half_rn(o[i] * inv);
      uint4* dst = reinterpret_cast<uint4*>(a.out + (long)(row0 + r) * a.ldo + 64 * h);
#pragma unroll
      for (int g = 0; g < 8; g++) dst[g] = reinterpret_cast<const uint4*>(out)[g];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(FA_TMEM_COLS) : "memory");
  }
}

bool attention_tc_supported(int ldq, int ldo, long T, long Tpad, int nkeys, const void* q, const void* K, const void* Vt,
                            const void* out) {
  return nkeys >= 1 && nkeys <= T && Tpad >= ((nkeys + FA_BK - 1) / FA_BK) * FA_BK && (Tpad % 64) == 0 &&
         (ldq % 8) == 0 && (ldo % 8) == 0 && ((uintptr_t)q % 16) == 0 && ((uintptr_t)K % 16) == 0 &&
         ((uintptr_t)Vt % 16) == 0 && ((uintptr_t)out % 16) == 0;
}

// q [n_slots*rows_per_slot][ldq] (head h at columns 64h..), K [n_slots][H][T][64], Vt [n_slots][H][64][Tpad] (zero beyond
// nkeys), every query of a slot attends to the first nkeys keys of that slot; out [rows][ldo] (head h at 64h..).
int launch_attention_tc(const __half* q, int ldq, const __half* K, const __half* Vt, long T, long Tpad, int H,
                        int rows_per_slot, int n_slots, int nkeys, float scale, __half* out, int ldo,
                        cudaStream_t st) {
  if (n_slots <= 0 || rows_per_slot <= 0) return 0;
  if (!attention_tc_supported(ldq, ldo, T, Tpad, nkeys, q, K, Vt, out)) {
    set_error("attention_tc: unsupported shape/alignment (nkeys=%d T=%ld Tpad=%ld ldq=%d ldo=%d)", nkeys, T, Tpad, ldq,
              ldo);
    return 1;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FaSmem) + 1024);
    attr_done = true;
  }
  CUtensorMap mq, mk, mv;
  const long rows = (long)n_slots * rows_per_slot;
  if (tc_make_map(&mq, q, rows, (long)H * 64, ldq, FA_BQ, 64) ||
      tc_make_map(&mk, K, (long)n_slots * H * T, 64, 64, FA_BK, 64) ||
      tc_make_map(&mv, Vt, (long)n_slots * H * 64, Tpad, Tpad, 64, 64))
    return 1;
  FaArgs a;
  a.out = out; a.ldo = ldo; a.H = H; a.rows_per_slot = rows_per_slot; a.nkeys = nkeys; a.T = T;
  a.sl2 = scale * 1.4426950408889634f;
  dim3 grid((rows_per_slot + FA_BQ - 1) / FA_BQ, H, n_slots);
  attention_tc_kernel<<<grid, FA_THREADS, sizeof(FaSmem) + 1024, st>>>(mq, mk, mv, a);
  count_launch();
  return check_launch("attention_tc_kernel") ? 0 : 1;
}

// V^T for the kernel above: dst[((slot*H + h)*64 + d)*Tpad + t] = src[m*ld + col0 + h*head_stride + d] (t < n), 0 beyond;
// m = slot*n + t.  One CTA per (64-key block, head, slot).
__global__ void __launch_bounds__(256)
    scatter_heads_t_kernel(const __half* __restrict__ src, int ld, int col0, int head_stride, int H, int n, long Tpad,
                           __half* __restrict__ dst) {
  __shared__ __half tile[64][72];  // [key][d], padded
  const int kb = blockIdx.x, h = blockIdx.y, slot = blockIdx.z, tid = threadIdx.x;
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const int idx = tid + 256 * it, key = idx >> 3, piece = idx & 7;
    const int t = kb * 64 + key;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (t < n) u = *reinterpret_cast<const uint4*>(src + ((long)slot * n + t) * ld + col0 + h * head_stride + 8 * piece);
    *reinterpret_cast<uint4*>(&tile[key][8 * piece]) = u;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const int idx = tid + 256 * it, d = idx >> 3, piece = idx & 7;
    __half v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = tile[8 * piece + i][d];
    *reinterpret_cast<uint4*>(dst + (((long)slot * H + h) * 64 + d) * Tpad + kb * 64 + 8 * piece) =
        *reinterpret_cast<const uint4*>(v);
  }
}

int launch_scatter_heads_t(const __half* src, int ld, int col0, int head_stride, int H, int n, long Tpad, int n_slots,
                           __half* dst, cudaStream_t st) {
  if (Tpad % 64) {
    set_error("scatter_heads_t: Tpad=%ld is not a multiple of 64", Tpad);
    return 1;
  }
  dim3 grid((unsigned)(Tpad / 64), H, n_slots);
  scatter_heads_t_kernel<<<grid, 256, 0, st>>>(src, ld, col0, head_stride, H, n, Tpad, dst);
  count_launch();
  return check_launch("scatter_heads_t_kernel") ? 0 : 1;
}

}  // namespace ma
