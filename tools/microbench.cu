// Micro-benchmarks behind the design of decode_mega.cu (run on the B200: nvcc -arch=sm_100a -O3 -o mb tools/microbench.cu && ./mb)
//  1. all-gather of a small vector through L2 with flagged 8-byte words (LL): every CTA writes its slice, every CTA
//     polls the whole vector -- the hand-off between two phases of the persistent decode kernel;
//  2. the same with a release/acquire counter barrier followed by plain ld.cg loads;
//  3. canonical GEMV from shared memory (22 rows x K=1024, 8 rows x K=4096), cycles per phase.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void ll_store(uint2* p, uint32_t d, uint32_t ep) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(d), "r"(ep) : "memory");
}
__device__ __forceinline__ uint4 ll_load2(const uint2* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ll_load2_relaxed(const uint2* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// words: vector of nwords flagged words; CTA c owns words [c*per, (c+1)*per)
template <int MODE>
__global__ void allgather_ll(uint2* words, int nwords, int per, int iters, unsigned long long* out) {
  const int tid = threadIdx.x, cta = blockIdx.x;
  unsigned long long t0 = 0;
  uint32_t acc = 0;
  for (int it = 1; it <= iters; it++) {
    if (it == 11 && cta == 0 && tid == 0) t0 = clock64();
    // produce: lane 0 of warps 0..per-1 writes one word each (like the GEMV epilogue)
    if (tid < per && cta * per + tid < nwords) ll_store(words + cta * per + tid, it * 7 + cta, it);
    // consume: every thread polls 2 words per round
    for (int u = tid; u < nwords / 2; u += blockDim.x) {
      uint4 v = MODE == 0 ? ll_load2(words + 2 * u) : ll_load2_relaxed(words + 2 * u);
      while ((int)(v.y - (uint32_t)it) < 0 || (int)(v.w - (uint32_t)it) < 0) v = MODE == 0 ? ll_load2(words + 2 * u) : ll_load2_relaxed(words + 2 * u);
      acc += v.x + v.z;
    }
    __syncthreads();
  }
  if (cta == 0 && tid == 0) { out[0] = clock64() - t0; out[1] = acc; }
}

// variant: only warp 0 polls (each lane a strided set of 16-byte units), others wait at the barrier
__global__ void allgather_ll_onewarp(uint2* words, int nwords, int per, int iters, unsigned long long* out) {
  const int tid = threadIdx.x, cta = blockIdx.x;
  unsigned long long t0 = 0;
  uint32_t acc = 0;
  for (int it = 1; it <= iters; it++) {
    if (it == 11 && cta == 0 && tid == 0) t0 = clock64();
    if (tid < per && cta * per + tid < nwords) ll_store(words + cta * per + tid, it * 7 + cta, it);
    if (tid < 32) {
      for (int u0 = 0; u0 < nwords / 2; u0 += 32 * 4) {   // 4 independent loads in flight per lane
        uint4 v[4];
        bool ok;
        do {
          ok = true;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int u = u0 + tid + 32 * k;
            if (u < nwords / 2) {
              v[k] = ll_load2(words + 2 * u);
              ok = ok && (int)(v[k].y - (uint32_t)it) >= 0 && (int)(v[k].w - (uint32_t)it) >= 0;
            }
          }
        } while (!ok);
#pragma unroll
        for (int k = 0; k < 4; k++) acc += v[k].x + v[k].z;
      }
    }
    __syncthreads();
  }
  if (cta == 0 && tid == 0) { out[0] = clock64() - t0; out[1] = acc; }
}

__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void allgather_barrier(uint32_t* vec, int nwords, int per, int iters, unsigned int* ctr, unsigned long long* out) {
  const int tid = threadIdx.x, cta = blockIdx.x;
  unsigned long long t0 = 0;
  uint32_t acc = 0;
  unsigned int target = 0;
  for (int it = 1; it <= iters; it++) {
    if (it == 11 && cta == 0 && tid == 0) t0 = clock64();
    if (tid < per && cta * per + tid < nwords) vec[cta * per + tid] = it * 7 + cta;
    __syncthreads();
    if (tid == 0) {
      target += gridDim.x;
      red_release_add(ctr, 1u);
      while ((int)(ld_acquire(ctr) - target) < 0) {}
    }
    __syncthreads();
    for (int u = tid; u < nwords / 4; u += blockDim.x) {
      uint4 v = __ldcg(reinterpret_cast<const uint4*>(vec) + u);
      acc += v.x + v.z;
    }
    __syncthreads();
  }
  if (cta == 0 && tid == 0) { out[0] = clock64() - t0; out[1] = acc; }
}

// single round trip: CTA 0 writes word, CTA 1 echoes, ...
__global__ void pingpong(uint2* w, int iters, unsigned long long* out) {
  const int cta = blockIdx.x;
  if (threadIdx.x != 0) return;
  unsigned long long t0 = clock64();
  for (int it = 1; it <= iters; it++) {
    if (cta == 0) {
      ll_store(w, it, it);
      uint4 v;
      do { v = ll_load2(w + 2); } while ((int)(v.y - (uint32_t)it) < 0);
    } else if (cta == gridDim.x - 1) {
      uint4 v;
      do { v = ll_load2(w); } while ((int)(v.y - (uint32_t)it) < 0);
      ll_store(w + 2, it, it);
    }
  }
  if (cta == 0) out[0] = clock64() - t0;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
template <int K>
__global__ void gemv_smem(int nrows, int iters, unsigned long long* out, float* sink) {
  extern __shared__ __align__(16) __half sm[];
  __half* sw = sm;               // [nrows][K]
  __half* xs = sm + (size_t)nrows * K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < nrows * K + K; i += blockDim.x) sm[i] = __float2half(0.001f * (i % 97));
  __syncthreads();
  unsigned long long t0 = clock64();
  float total = 0.f;
  for (int it = 0; it < iters; it++) {
    const int npairs = nrows / 2;
    for (int i = 0; i < 4; i += 2) {
      const int pA = warp + 8 * i, pB = pA + 8;
      if (pA >= npairs) break;
      const bool hasB = pB < npairs;
      const __half* wA = sw + (size_t)(2 * pA) * K + 8 * lane;
      const __half* wB = sw + (size_t)(2 * (hasB ? pB : pA)) * K + 8 * lane;
      float acc[4] = {0, 0, 0, 0};
#pragma unroll 4
      for (int g = 0; g < K / 256; g++) {
        float xf[8], f[4][8];
        unpack8(*reinterpret_cast<const uint4*>(xs + 256 * g + 8 * lane), xf);
        unpack8(*reinterpret_cast<const uint4*>(wA + 256 * g), f[0]);
        unpack8(*reinterpret_cast<const uint4*>(wA + K + 256 * g), f[1]);
        unpack8(*reinterpret_cast<const uint4*>(wB + 256 * g), f[2]);
        unpack8(*reinterpret_cast<const uint4*>(wB + K + 256 * g), f[3]);
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
          for (int r = 0; r < 4; r++) acc[r] = __fmaf_rn(f[r][j], xf[j], acc[r]);
      }
      float k = acc[0] + acc[1] + acc[2] + acc[3];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
      total += k;
    }
    __syncthreads();
  }
  if (tid == 0) { out[0] = clock64() - t0; }
  if (total == 123.f) sink[0] = total;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("SMs %d, clock %d kHz\n", sms, khz);
  const int grid = 147;
  uint2* words; unsigned long long* out; unsigned int* ctr; float* sink;
  cudaMalloc(&words, 1 << 20); cudaMalloc(&out, 64); cudaMalloc(&ctr, 64); cudaMalloc(&sink, 64);
  unsigned long long h[2];
  const int iters = 2010;
  auto report = [&](const char* name, int n) {
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-44s %8.1f cycles/iter = %6.3f us  (%s)\n", name, (double)h[0] / n, (double)h[0] / n / (khz / 1000.0), cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
  };
  for (int rep = 0; rep < 2; rep++) {
    cudaMemset(words, 0, 1 << 20);
    allgather_ll<0><<<grid, 256>>>(words, 512, 4, iters, out);      // 1024 halfs (yb/ya/attn): 4 words per CTA (128 producers)
    report("LL all-gather 512 words (volatile)", iters - 10);
    cudaMemset(words, 0, 1 << 20);
    allgather_ll<1><<<grid, 256>>>(words, 512, 4, iters, out);
    report("LL all-gather 512 words (relaxed.gpu)", iters - 10);
    cudaMemset(words, 0, 1 << 20);
    allgather_ll<0><<<grid, 256>>>(words, 2048, 14, iters, out);    // 4096 halfs (fc1 output)
    report("LL all-gather 2048 words (volatile)", iters - 10);
    cudaMemset(words, 0, 1 << 20);
    allgather_ll_onewarp<<<grid, 256>>>(words, 512, 4, iters, out);
    report("LL all-gather 512 words, one polling warp", iters - 10);
    cudaMemset(words, 0, 1 << 20);
    allgather_ll_onewarp<<<grid, 256>>>(words, 2048, 14, iters, out);
    report("LL all-gather 2048 words, one polling warp", iters - 10);
    cudaMemset(words, 0, 1 << 20);
    allgather_ll<0><<<16, 256>>>(words, 512, 32, iters, out);
    report("LL all-gather 512 words, 16 CTAs only", iters - 10);
    cudaMemset(words, 0, 1 << 20); cudaMemset(ctr, 0, 64);
    allgather_barrier<<<grid, 256>>>((uint32_t*)words, 512, 4, iters, ctr, out);
    report("barrier + ld.cg all-gather 512 words", iters - 10);
    cudaMemset(words, 0, 1 << 20);
    pingpong<<<grid, 32>>>(words, 2000, out);
    report("ping-pong CTA0 <-> CTA146 (2 hand-offs)", 2000);
    cudaMemset(words, 0, 1 << 20);
    pingpong<<<2, 32>>>(words, 2000, out);
    report("ping-pong CTA0 <-> CTA1 (2 hand-offs)", 2000);
  }
  cudaFuncSetAttribute(gemv_smem<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  cudaFuncSetAttribute(gemv_smem<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  gemv_smem<1024><<<1, 256, (22 * 1024 + 1024) * 2>>>(22, 1000, out, sink); report("gemv 22 rows K=1024 (1 CTA)", 1000);
  gemv_smem<1024><<<1, 256, (28 * 1024 + 1024) * 2>>>(28, 1000, out, sink); report("gemv 28 rows K=1024 (1 CTA)", 1000);
  gemv_smem<1024><<<1, 256, (8 * 1024 + 1024) * 2>>>(8, 1000, out, sink);   report("gemv 8 rows K=1024 (1 CTA)", 1000);
  gemv_smem<4096><<<1, 256, (8 * 4096 + 4096) * 2>>>(8, 1000, out, sink);   report("gemv 8 rows K=4096 (1 CTA)", 1000);
  return 0;
}
