// surface.cu -- mesh -> point cloud on the GPU: area-weighted surface sampling with face normals.
//
// Replaces `trimesh.Trimesh.sample(count, return_index=True)` + `mesh.face_normals[idx]` of the reference's
// pre-processing (/root/reference/mesh_to_pc.py:42-57): a face is drawn with probability proportional to its area
// (inverse CDF over the cumulative areas), a point uniformly inside it (two uniforms, reflected into the triangle --
// what trimesh.sample.sample_surface does), and the face's unit normal is appended.  Output fp16 [n][6], the dtype the
// reference feeds the encoder (np.float16, mesh_to_pc.py:53).  The random stream is Philox4x32-10 keyed by
// (seed, sample), not numpy's Mersenne twister: same distribution, different individual points.
#include "canon.cuh"
#include "internal.h"

namespace ma {

__device__ __forceinline__ uint32_t sf_mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
  const unsigned long long p = (unsigned long long)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
// three uniforms in [0,1) for sample i
__device__ __forceinline__ float3 sf_philox3(unsigned long long seed, uint32_t i) {
  uint32_t c0 = i, c1 = 0x53555246u, c2 = 0x4d455348u, c3 = 0x414e5954u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = sf_mulhilo(0xD2511F53u, c0, &hi0), lo1 = sf_mulhilo(0xCD9E8D57u, c2, &hi1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float s = 1.0f / 16777216.0f;
  return make_float3((float)(c0 >> 8) * s, (float)(c1 >> 8) * s, (float)(c2 >> 8) * s);
}

__device__ __forceinline__ float3 sf_vertex(const float* v, int i) { return make_float3(v[3 * i], v[3 * i + 1], v[3 * i + 2]); }

// area[f] (double: the cumulative sum of up to millions of faces must stay monotone and exact enough for the search)
__global__ void surface_area_kernel(const float* __restrict__ v, const int32_t* __restrict__ faces, int F, double* __restrict__ area) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float3 a = sf_vertex(v, faces[3 * f]), b = sf_vertex(v, faces[3 * f + 1]), c = sf_vertex(v, faces[3 * f + 2]);
  const double ux = (double)b.x - a.x, uy = (double)b.y - a.y, uz = (double)b.z - a.z;
  const double wx = (double)c.x - a.x, wy = (double)c.y - a.y, wz = (double)c.z - a.z;
  const double nx = uy * wz - uz * wy, ny = uz * wx - ux * wz, nz = ux * wy - uy * wx;
  area[f] = 0.5 * sqrt(nx * nx + ny * ny + nz * nz);
}

// in-place inclusive scan by ONE CTA of 1024 threads walking the array in tiles (F is at most a few million)
__global__ void __launch_bounds__(1024) surface_scan_kernel(double* __restrict__ a, int F) {
  __shared__ double wsum[32];
  __shared__ double carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0.0;
  __syncthreads();
  for (int base = 0; base < F; base += 1024) {
    const int i = base + tid;
    double x = i < F ? a[i] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      double s = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const double off = carry + (warp ? wsum[warp - 1] : 0.0);
    if (i < F) a[i] = x + off;
    __syncthreads();
    if (tid == 1023) carry = x + off;
    __syncthreads();
  }
}

__global__ void surface_sample_kernel(const float* __restrict__ v, const int32_t* __restrict__ faces, int F,
                                      const double* __restrict__ cum, int n, unsigned long long seed,
                                      __half* __restrict__ out, int32_t* __restrict__ face_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float3 u = sf_philox3(seed, (uint32_t)i);
  const double target = (double)u.x * cum[F - 1];
  int lo = 0, hi = F - 1;                       // first face whose cumulative area exceeds the target
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cum[mid] > target) hi = mid; else lo = mid + 1;
  }
  const int f = lo;
  const float3 a = sf_vertex(v, faces[3 * f]), b = sf_vertex(v, faces[3 * f + 1]), c = sf_vertex(v, faces[3 * f + 2]);
  float r1 = u.y, r2 = u.z;
  if (r1 + r2 > 1.0f) { r1 = 1.0f - r1; r2 = 1.0f - r2; }
  const float ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z, wx = c.x - a.x, wy = c.y - a.y, wz = c.z - a.z;
  const float px = a.x + r1 * ux + r2 * wx, py = a.y + r1 * uy + r2 * wy, pz = a.z + r1 * uz + r2 * wz;
  float nx = uy * wz - uz * wy, ny = uz * wx - ux * wz, nz = ux * wy - uy * wx;
  const float ln = sqrtf(nx * nx + ny * ny + nz * nz);
  const float inv = ln > 0.0f ? 1.0f / ln : 1.0f;
  __half* o = out + (size_t)i * 6;
  o[0] = __float2half_rn(px); o[1] = __float2half_rn(py); o[2] = __float2half_rn(pz);
  o[3] = __float2half_rn(nx * inv); o[4] = __float2half_rn(ny * inv); o[5] = __float2half_rn(nz * inv);
  if (face_idx) face_idx[i] = f;
}

}  // namespace ma

using namespace ma;

extern "C" {

size_t ma_sample_surface_workspace_bytes(int n_faces) { return (size_t)(n_faces > 0 ? n_faces : 1) * sizeof(double) + 256; }

int ma_sample_surface(const float* vertices, const int32_t* faces, int n_faces, int n_samples, unsigned long long seed,
                      void* out_pc_normal, int32_t* out_face_idx, void* ws, void* stream) {
  if (!vertices || !faces || !out_pc_normal || !ws || n_faces <= 0 || n_samples <= 0) {
    set_error("ma_sample_surface: bad arguments");
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  double* cum = reinterpret_cast<double*>(ws);
  surface_area_kernel<<<(n_faces + 255) / 256, 256, 0, st>>>(vertices, faces, n_faces, cum);
  surface_scan_kernel<<<1, 1024, 0, st>>>(cum, n_faces);
  surface_sample_kernel<<<(n_samples + 255) / 256, 256, 0, st>>>(vertices, faces, n_faces, cum, n_samples, seed,
                                                                  reinterpret_cast<__half*>(out_pc_normal), out_face_idx);
  count_launch(3);
  return check_launch("ma_sample_surface") ? 0 : 1;
}

}  // extern "C"
