// Micro-benchmarks behind the cluster design of decode_mega.cu (round 2).  Run on the B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mbc tools/microbench_cluster.cu && tools/mbc
//  1. how many 8-CTA clusters of the persistent kernel's shape (512 threads, 227 KB shared memory) are co-resident;
//  2. fma.rn.f32.f16 (SASS FHFMA) against convert + FFMA on random fp16 data incl. subnormals: bit equality;
//  3. DSMEM all-gather inside a cluster with st.async (data + mbarrier complete_tx in one instruction);
//  4. the full chip-wide hand-off of the new design: every CTA publishes 128 flagged fp32 words (LL) -> the CTAs of
//     equal rank in the 16 clusters reduce their slice -> st.async all-gather of the reduced slice inside the cluster;
//  5. barrier.cluster arrive + wait, for reference.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try(bar, parity)) {} }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void ll_store(uint2* p, uint32_t d, uint32_t ep) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(d), "r"(ep) : "memory");
}
__device__ __forceinline__ uint2 ll_load1(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

// ---- 1. occupancy probe -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) shape_kernel(int* out) {
  extern __shared__ unsigned char sm[];
  if (threadIdx.x == 0 && out) out[blockIdx.x] = sm[0];
}

// ---- 2. FHFMA equivalence ---------------------------------------------------------------------------------------
__global__ void fhfma_check(const unsigned short* a, const unsigned short* b, int n, int chain, unsigned int* mism, float* sink) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float c0 = 0.0f, c1 = 0.0f;
  for (int k = 0; k < chain; k++) {
    const unsigned short x = a[(i + 131 * k) % n], y = b[(i * 7 + 17 * k) % n];
    c0 = __fmaf_rn(__half2float(__ushort_as_half(x)), __half2float(__ushort_as_half(y)), c0);
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c1) : "h"(x), "h"(y));
    if (__float_as_uint(c0) != __float_as_uint(c1)) atomicAdd(mism, 1u);
    if (isinf(c0) || isnan(c0)) { c0 = 0.0f; c1 = 0.0f; }
  }
  sink[i] = c0 + c1;
}

// ---- 3. st.async all-gather in a cluster ------------------------------------------------------------------------
// every thread sends `per_thread` 4-byte words to each... thread t sends word (t >> 3) to rank (t & 7): 64 words x 8
template <int NWORDS_PER_CTA>   // words each CTA contributes (sent to all 8 CTAs)
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(512, 1) dsmem_allgather(int iters, unsigned long long* out) {
  __shared__ uint64_t bar;
  __shared__ uint32_t buf[8 * NWORDS_PER_CTA];
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_rank();
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  cluster_sync_all();
  unsigned long long t0 = 0;
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    if (it == 10 && tid == 0) t0 = clock64();
    if (tid == 0) mbar_expect_tx(&bar, 8 * NWORDS_PER_CTA * 4);
    if (tid < NWORDS_PER_CTA * 8) {
      const uint32_t dst = tid & 7, w = tid >> 3;
      st_async_b32(mapa(smem_u32(&buf[rank * NWORDS_PER_CTA + w]), dst), it * 3 + w + acc, mapa(smem_u32(&bar), dst));
    }
    mbar_wait(&bar, it & 1);
    acc += buf[(tid * 5) % (8 * NWORDS_PER_CTA)];
    __syncthreads();   // everyone has read before the next round may overwrite (also ordered by the all-to-all itself)
  }
  if (tid == 0 && blockIdx.x == 0) { out[0] = clock64() - t0; out[1] = acc; }
  cluster_sync_all();
}

// ---- 4. publish (LL) -> slice reduce across the 16 clusters -> st.async all-gather -------------------------------
// part[rank j][cluster h][row 0..127] flagged words
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(512, 1) exchange_full(uint2* part, int iters, unsigned long long* out, int skew) {
  __shared__ uint64_t bar;
  __shared__ uint32_t ybuf[512];
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t j = cluster_rank(), h = cluster_id();
  const int ncl = gridDim.x / 8;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  cluster_sync_all();
  unsigned long long t0 = 0;
  float acc = 0.0f;
  for (int it = 1; it <= iters; it++) {
    if (it == 11 && tid == 0) t0 = clock64();
    if (skew && ((blockIdx.x * 7 + it) % 13) == 0) { const long long s = clock64(); while (clock64() - s < 600) {} }
    if (tid == 0) mbar_expect_tx(&bar, 8 * 64 * 4);
    // publish: lanes with (lane & 3) == 0 of every warp write one word: 8 per warp x 16 warps = 128 rows
    if ((lane & 3) == 0) {
      const int row = (tid >> 5) * 8 + (lane >> 2);
      ll_store(part + ((size_t)j * ncl + h) * 128 + row, __float_as_uint(1.0f + acc * 1e-9f + row), it);
    }
    // slice reduce: thread (row = tid >> 2, q = tid & 3) polls 4 words
    const int row = tid >> 2, q = tid & 3;
    float v[4];
    {
      uint2 w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) w[i] = ll_load1(part + ((size_t)j * ncl + ((4 * q + i) % ncl)) * 128 + row);
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (w[i].y != (uint32_t)it) { ok = false; w[i] = ll_load1(part + ((size_t)j * ncl + ((4 * q + i) % ncl)) * 128 + row); }
        if (ok) break;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = __uint_as_float(w[i].x);
    }
    float s = __fadd_rn(__fadd_rn(v[0], v[1]), __fadd_rn(v[2], v[3]));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
    const __half y = __float2half_rn(s);
    const __half y2 = __shfl_xor_sync(0xffffffffu, y, 4);
    const uint32_t word = (lane & 4) ? ((uint32_t)__half_as_ushort(y2) | ((uint32_t)__half_as_ushort(y) << 16))
                                     : ((uint32_t)__half_as_ushort(y) | ((uint32_t)__half_as_ushort(y2) << 16));
    st_async_b32(mapa(smem_u32(&ybuf[j * 64 + (tid >> 3)]), tid & 7), word, mapa(smem_u32(&bar), tid & 7));
    mbar_wait(&bar, (it - 1) & 1);
    acc += __half2float(__ushort_as_half((unsigned short)(ybuf[tid] & 0xffff)));
    __syncthreads();
  }
  if (tid == 0 && blockIdx.x == 0) { out[0] = clock64() - t0; out[1] = (unsigned long long)acc; }
  cluster_sync_all();
}

// ---- 5. cluster barrier ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(512, 1) cluster_barrier_loop(int iters, unsigned long long* out) {
  unsigned long long t0 = 0;
  for (int it = 0; it < iters; it++) {
    if (it == 10 && threadIdx.x == 0) t0 = clock64();
    cluster_sync_all();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}


// ---- 6. order-free reduction in L2: every CTA adds its partial of every row with ONE 64-bit atomic that carries the
// fixed-point value (low 48 bits, two's complement) and a contribution count (bits 48..63); every CTA polls the 4
// rows it needs until the count is complete.  acc has 3 rotating buffers (the one used next is zeroed by its owner).
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
template <int ROWS_PER_CTA>   // 1024: every CTA contributes to every row (fc2); 128: CTA (g, j) to rows 128 j.. (out_proj)
__global__ void __launch_bounds__(512, 1) exchange_atomic(unsigned long long* acc, int iters, unsigned long long* out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
  const int ncontrib = ROWS_PER_CTA == 1024 ? gridDim.x : gridDim.x / 8;
  unsigned long long t0 = 0;
  float accf = 0.0f;
  for (int it = 0; it < iters; it++) {
    if (it == 10 && tid == 0) t0 = clock64();
    unsigned long long* a = acc + (size_t)(it % 3) * 1024;
    unsigned long long* z = acc + (size_t)((it + 1) % 3) * 1024;
    if (tid < 8) z[(cta % 128) * 8 + tid] = 0ull;   // zero the buffer of the next use (owner: 8 rows per CTA)
    // contributions: one lane in four holds a row's partial (8 rows per warp instruction)
    for (int r0 = 0; r0 < ROWS_PER_CTA; r0 += 128) {
      const int r = r0 + warp * 8 + (lane >> 2);
      const int row = ROWS_PER_CTA == 1024 ? r : (cta & 7) * 128 + r;
      const long long fix = __float2ll_rn((float)((row * 7 + cta + it) % 13 - 6 + accf * 1e-20f) * 268435456.0f);
      if ((lane & 3) == 0) red_add_u64(a + row, (1ull << 48) + (unsigned long long)fix);
    }
    // read: thread t < 256 needs rows 4t..4t+3
    if (tid < 256) {
      unsigned long long w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) w[i] = ld_volatile_u64(a + 4 * tid + i);
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 4; i++)
          if ((int)((w[i] + (1ull << 47)) >> 48) != ncontrib) { ok = false; w[i] = ld_volatile_u64(a + 4 * tid + i); }
        if (ok) break;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) accf += (float)(long long)(w[i] - ((unsigned long long)ncontrib << 48)) * 3.7252903e-9f;
    }
    __syncthreads();
  }
  if (tid == 0 && cta == 0) { out[0] = clock64() - t0; out[1] = (unsigned long long)accf; }
}

// compile check only: L2 prefetch of a contiguous range
__global__ void l2_prefetch_probe(const void* p, unsigned bytes) {
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

static unsigned long long* g_out;
static double g_mhz;
static void report(const char* what, int iters) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s FAILED: %s\n", what, cudaGetErrorString(e)); return; }
  unsigned long long h[2];
  cudaMemcpy(h, g_out, 16, cudaMemcpyDeviceToHost);
  printf("%-58s %9.1f cycles/iter = %6.3f us\n", what, (double)h[0] / iters, (double)h[0] / iters / g_mhz);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  g_mhz = clk / 1000.0;
  printf("%s: %d SMs, %d kHz\n", prop.name, prop.multiProcessorCount, clk);
  cudaMalloc(&g_out, 64);

  // 1. co-resident clusters of the persistent kernel's shape
  for (int smem : {227 * 1024, 200 * 1024, 100 * 1024}) {
    cudaFuncSetAttribute(shape_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(shape_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {2, 4, 8, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, shape_kernel, &cfg);
      printf("max active clusters: smem %3d KB, cluster %2d -> %d clusters = %d CTAs (%s)\n", smem / 1024, cs, n, n * cs,
             e == cudaSuccess ? "ok" : cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  {  // cooperative + cluster launch accepted?
    const int smem = 227 * 1024;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    int* nul = nullptr;
    cudaError_t e = cudaLaunchKernelEx(&cfg, shape_kernel, nul);
    cudaError_t e2 = cudaDeviceSynchronize();
    printf("cluster(8) + cooperative launch of 128 CTAs x 227 KB: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaGetLastError();
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, shape_kernel, nul);
    e2 = cudaDeviceSynchronize();
    printf("cluster(8) launch of 128 CTAs x 227 KB (not cooperative): %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaGetLastError();
  }

  // 2. FHFMA
  {
    const int n = 1 << 20;
    unsigned short *ha = (unsigned short*)malloc(n * 2), *hb = (unsigned short*)malloc(n * 2);
    srand(1);
    for (int i = 0; i < n; i++) {
      // mix: arbitrary bit patterns (incl. subnormals, inf, nan), small weights, activations
      const int kind = rand() % 4;
      if (kind == 0) { ha[i] = (unsigned short)rand(); hb[i] = (unsigned short)rand(); }
      else if (kind == 1) { ha[i] = (unsigned short)(rand() % 0x0400) | (rand() & 1 ? 0x8000 : 0); hb[i] = (unsigned short)rand(); }  // subnormal a
      else {
        const float x = ((rand() % 20001) - 10000) * 2e-6f * (kind == 2 ? 1.0f : 100.0f), y = ((rand() % 20001) - 10000) * 3e-4f;
        ha[i] = __half_as_ushort(__float2half_rn(x)); hb[i] = __half_as_ushort(__float2half_rn(y));
      }
    }
    unsigned short *da, *db; unsigned int* dm; float* ds;
    cudaMalloc(&da, n * 2); cudaMalloc(&db, n * 2); cudaMalloc(&dm, 4); cudaMalloc(&ds, n * 4);
    cudaMemcpy(da, ha, n * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 2, cudaMemcpyHostToDevice);
    cudaMemset(dm, 0, 4);
    fhfma_check<<<n / 256, 256>>>(da, db, n, 64, dm, ds);
    unsigned int mm = 0;
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&mm, dm, 4, cudaMemcpyDeviceToHost);
    printf("fma.rn.f32.f16 vs cvt+ffma: %u mismatching results out of %lld (%s)\n", mm, (long long)n * 64, cudaGetErrorString(e));
  }

  const int iters = 2010;
  dsmem_allgather<64><<<128, 512>>>(iters, g_out);  report("st.async all-gather 8 x 256 B (512 stores/CTA), 16 clusters", iters - 10);
  dsmem_allgather<12><<<128, 512>>>(iters, g_out);  report("st.async all-gather 8 x 48 B (96 stores/CTA), 16 clusters", iters - 10);
  dsmem_allgather<64><<<8, 512>>>(iters, g_out);    report("st.async all-gather 8 x 256 B, 1 cluster", iters - 10);
  cluster_barrier_loop<<<128, 512>>>(iters, g_out); report("barrier.cluster arrive+wait, 16 clusters", iters - 10);
  uint2* part;
  cudaMalloc(&part, 8 * 16 * 128 * 8);
  cudaMemset(part, 0, 8 * 16 * 128 * 8);
  exchange_full<<<128, 512>>>(part, iters, g_out, 0);  report("LL publish + slice reduce(16) + st.async all-gather", iters - 10);
  cudaMemset(part, 0, 8 * 16 * 128 * 8);
  exchange_full<<<128, 512>>>(part, iters, g_out, 1);  report("  same with injected skew (0.3 us on 1/13 of the CTAs)", iters - 10);
  cudaMemset(part, 0, 8 * 16 * 128 * 8);
  exchange_full<<<8, 512>>>(part, iters, g_out, 0);    report("  same, 1 cluster only (reduce over 1)", iters - 10);
  unsigned long long* acc;
  cudaMalloc(&acc, 3 * 1024 * 8);
  cudaMemset(acc, 0, 3 * 1024 * 8);
  {
    void* args[] = {&acc, (void*)&iters, &g_out};
    cudaLaunchCooperativeKernel((void*)exchange_atomic<1024>, dim3(128), dim3(512), args, 0, 0);
    report("atomic fixed-point reduce: 128 CTAs x 1024 rows (fc2 shape)", iters - 10);
    cudaMemset(acc, 0, 3 * 1024 * 8);
    cudaLaunchCooperativeKernel((void*)exchange_atomic<128>, dim3(128), dim3(512), args, 0, 0);
    report("atomic fixed-point reduce: 16 contributions per row (out_proj shape)", iters - 10);
    cudaMemset(acc, 0, 3 * 1024 * 8);
    cudaLaunchCooperativeKernel((void*)exchange_atomic<1024>, dim3(144), dim3(512), args, 0, 0);
    report("atomic fixed-point reduce: 144 CTAs x 1024 rows", iters - 10);
  }
  return 0;
}
