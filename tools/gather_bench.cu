// Random row-gather ceiling of one B200: how fast can HBM deliver independent random rows of a given size?
// (Development aid; the traversal kernel's roofline denominator stays the measured copy bandwidth, this number
// says how much of the gap is the access pattern.)   nvcc -arch=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// variant 0: LDG.128 per lane; a warp reads UNR rows (row_bytes = 512*CH) back to back, then reduces
template <int CH, int UNR>
__global__ void gather_ldg(const uint4 *rows, uint32_t n_rows, uint32_t per_warp, float *out) {
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  uint32_t s = mix(gw * 2654435761u + 1u);
  for (uint32_t it = 0; it < per_warp; it += UNR) {
    uint4 v[UNR][CH];
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      s = mix(s + 0x9e3779b9u);
      const uint32_t id = s % n_rows;
#pragma unroll
      for (int c = 0; c < CH; c++) v[u][c] = __ldg(rows + (size_t)id * (32 * CH) + c * 32 + lane);
    }
#pragma unroll
    for (int u = 0; u < UNR; u++)
#pragma unroll
      for (int c = 0; c < CH; c++) acc += __uint_as_float(v[u][c].x ^ v[u][c].y ^ v[u][c].z ^ v[u][c].w);
  }
  if (acc == 12345.678f) out[gw] = acc;
}
// variant 1: sub-warp groups of G lanes read rows of G*16 bytes (G = 8: 128 B, 16: 256 B)
template <int G, int UNR>
__global__ void gather_small(const uint4 *rows, uint32_t n_rows, uint32_t per_grp, float *out) {
  const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x, grp = gt / G, l = gt % G;
  float acc = 0.f;
  uint32_t s = mix(grp * 2654435761u + 1u);
  for (uint32_t it = 0; it < per_grp; it += UNR) {
    uint4 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) { s = mix(s + 0x9e3779b9u); v[u] = __ldg(rows + (size_t)(s % n_rows) * G + l); }
#pragma unroll
    for (int u = 0; u < UNR; u++) acc += __uint_as_float(v[u].x ^ v[u].y ^ v[u].z ^ v[u].w);
  }
  if (acc == 12345.678f) out[gt] = acc;
}
// variant 2: cp.async 16 B per lane into shared memory, UNR rows of 512 B in flight per warp
template <int UNR>
__global__ void gather_cpasync(const uint8_t *rows, uint32_t n_rows, uint32_t per_warp, float *out) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + warp * UNR * 512 + lane * 16;
  float acc = 0.f;
  uint32_t s = mix(gw * 2654435761u + 1u);
  for (uint32_t it = 0; it < per_warp; it += UNR) {
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      s = mix(s + 0x9e3779b9u);
      const uint8_t *src = rows + (size_t)(s % n_rows) * 512 + lane * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + u * 512), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      uint4 v;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + u * 512));
      acc += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w);
    }
    __syncwarp();
  }
  if (acc == 12345.678f) out[gw] = acc;
}
// streaming read for comparison
__global__ void stream_read(const uint4 *p, size_t n16, float *out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = __ldg(p + i);
    acc += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w);
  }
  if (acc == 12345.678f) out[0] = acc;
}

template <typename F> static float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; i++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main(int argc, char **argv) {
  const size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 0.5) * (1ull << 30);   // table size in GiB (0.5 = the 1M x 128 f32 set)
  uint8_t *d; float *out;
  CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 1, bytes)); CK(cudaMalloc(&out, 64 << 20));
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("table %.2f GiB, %d SMs\n", bytes / double(1ull << 30), sms);
  { float ms = timeit([&] { stream_read<<<sms * 16, 512>>>((const uint4 *)d, bytes / 16, out); });
    printf("stream read            : %7.1f GB/s\n", bytes / ms / 1e6); }
  const uint32_t total_rows_512 = 40u << 20;   // 20 GiB of traffic per launch at 512 B
#define RUN_LDG(CH, UNR, WPS)                                                                                    \
  { const uint32_t warps = sms * WPS, n_rows = bytes / (512 * CH), per = (total_rows_512 / CH / warps) / UNR * UNR; \
    float ms = timeit([&] { gather_ldg<CH, UNR><<<warps / 4, 128>>>((const uint4 *)d, n_rows, per, out); });    \
    printf("ldg  row %5d B unr %2d warps/SM %2d : %7.1f GB/s\n", 512 * CH, UNR, WPS, (double)per * warps * 512 * CH / ms / 1e6); }
  RUN_LDG(1, 4, 32) RUN_LDG(1, 8, 32) RUN_LDG(1, 16, 32) RUN_LDG(1, 8, 64) RUN_LDG(1, 16, 48) RUN_LDG(1, 8, 16)
  RUN_LDG(2, 4, 32) RUN_LDG(2, 8, 32) RUN_LDG(4, 4, 32) RUN_LDG(8, 2, 32)
#define RUN_SMALL(G, UNR, WPS)                                                                                   \
  { const uint32_t thr = sms * WPS * 32, n_rows = bytes / (16 * G), per = (uint32_t)(((size_t)total_rows_512 * 32 / 2) / thr) / UNR * UNR; \
    float ms = timeit([&] { gather_small<G, UNR><<<thr / 128, 128>>>((const uint4 *)d, n_rows, per, out); });   \
    printf("ldg  row %5d B unr %2d warps/SM %2d : %7.1f GB/s\n", 16 * G, UNR, WPS, (double)per * thr * 16 / ms / 1e6); }
  RUN_SMALL(2, 8, 32) RUN_SMALL(4, 8, 32) RUN_SMALL(8, 8, 32) RUN_SMALL(16, 8, 32) RUN_SMALL(8, 16, 64)
#define RUN_CPA(UNR, WPS)                                                                                        \
  { const uint32_t warps = sms * WPS, n_rows = bytes / 512, per = (total_rows_512 / warps) / UNR * UNR;          \
    CK(cudaFuncSetAttribute(gather_cpasync<UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * UNR * 512));   \
    float ms = timeit([&] { gather_cpasync<UNR><<<warps / 4, 128, 4 * UNR * 512>>>(d, n_rows, per, out); });     \
    printf("cp.async row 512 B unr %2d warps/SM %2d : %7.1f GB/s\n", UNR, WPS, (double)per * warps * 512 / ms / 1e6); }
  RUN_CPA(8, 32) RUN_CPA(16, 32) RUN_CPA(16, 16) RUN_CPA(32, 16) RUN_CPA(8, 48)
  return 0;
}
