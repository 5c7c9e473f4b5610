// gather_ceiling2.cu — follow-up diagnostic: WHY does a 64-byte row read cost 128 bytes of DRAM traffic?
// (ncu on the fused gather, 10 GB tables: dram read 2.77 GB = 20.4 M rows x 128 B + ids; see
// profiles/r1_gather_ceiling.md).  Sweeps the load flavour and cudaLimitMaxL2FetchGranularity for
// 64-byte rows.  Build like gather_ceiling.cu.  Usage: gather_ceiling2 [reps]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t) x;
}
__global__ void fill_kernel(float4* p, int64_t n) {
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
    p[i] = make_float4((float) (i & 255), 1.f, 2.f, 3.f);
}
__global__ void ids_kernel(int32_t* ids, int64_t n, uint32_t rows) {
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
    ids[i] = (int32_t) (mix((uint64_t) i * 2654435761ull + 12345) % rows);
}
enum { LD_DEFAULT = 0, LD_CG = 1, LD_NC_NOALLOC = 2, LD_CS = 3, LD_NC_L2_64 = 4, LD_RELAXED = 5 };
template <int MODE>
__device__ __forceinline__ float4 ld16(const float4* p) {
  float4 r;
  if (MODE == LD_DEFAULT) return __ldg(p);
  if (MODE == LD_CG) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == LD_NC_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == LD_CS) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == LD_NC_L2_64) asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == LD_RELAXED) asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
template <int MODE, int U, bool WRITE>
__global__ void __launch_bounds__(256) row64(const float4* __restrict__ table, const int32_t* __restrict__ ids,
                                             float4* __restrict__ out, int64_t nitems, float* sink) {
  const int sub = threadIdx.x & 3;
  const int64_t ngroups = ((int64_t) gridDim.x * blockDim.x) >> 2;
  const int64_t group = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  float acc = 0.f;
  for (int64_t base = group; base < nitems; base += ngroups * U) {
    int64_t row[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t it = base + (int64_t) u * ngroups; row[u] = it < nitems ? ids[it] : -1; }
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = row[u] >= 0 ? ld16<MODE>(table + row[u] * 4 + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t it = base + (int64_t) u * ngroups;
      if (WRITE) { if (it < nitems) out[it * 4 + sub] = v[u]; } else acc += v[u].x + v[u].w;
    }
  }
  if (!WRITE && acc == 123456.789f) *sink = acc;
}
static int g_reps = 5;
template <typename F> static float time_ms(F launch) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a)); for (int i = 0; i < g_reps; ++i) launch(); CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b)); CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms / g_reps;
}
template <int MODE, bool WRITE>
static void run(const char* name, int gran, const float4* table, const int32_t* ids, float4* out, int64_t n, float* sink) {
  const float ms = time_ms([&] { row64<MODE, 8, WRITE><<<148 * 16, 256>>>(table, ids, out, n, sink); });
  printf("{\"l2_fetch\": %d, \"load\": \"%s\", \"kernel\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f}\n", gran, name,
         WRITE ? "row_copy" : "row_read", ms, (double) n * (4.0 + 64 + (WRITE ? 64 : 0)) / ms / 1e6);
  fflush(stdout);
}
int main(int argc, char** argv) {
  if (argc > 1) g_reps = atoi(argv[1]);
  const int64_t nitems = 20447232, table_f4 = 625000000;
  float4 *table, *out; int32_t* ids; float* sink;
  CK(cudaMalloc(&table, table_f4 * 16)); CK(cudaMalloc(&out, nitems * 64)); CK(cudaMalloc(&ids, nitems * 4)); CK(cudaMalloc(&sink, 4));
  fill_kernel<<<148 * 8, 256>>>(table, table_f4);
  ids_kernel<<<148 * 8, 256>>>(ids, nitems, (uint32_t) (table_f4 / 4));
  CK(cudaDeviceSynchronize());
  size_t g0 = 0; CK(cudaDeviceGetLimit(&g0, cudaLimitMaxL2FetchGranularity));
  printf("{\"default_l2_fetch_granularity\": %zu}\n", g0);
  const int grans[4] = {(int) g0, 32, 64, 128};
  for (int gi = 0; gi < 4; ++gi) {
    const int g = grans[gi];
    if (gi > 0) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t) g);
      size_t now = 0; cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity);
      printf("{\"set_l2_fetch\": %d, \"rc\": \"%s\", \"now\": %zu}\n", g, cudaGetErrorString(e), now);
      if (e != cudaSuccess) { cudaGetLastError(); continue; }
    }
    run<LD_DEFAULT, true>("ldg", g, table, ids, out, nitems, sink);
    run<LD_DEFAULT, false>("ldg", g, table, ids, out, nitems, sink);
    run<LD_CG, true>("cg", g, table, ids, out, nitems, sink);
    run<LD_CG, false>("cg", g, table, ids, out, nitems, sink);
    run<LD_NC_NOALLOC, true>("nc.no_allocate", g, table, ids, out, nitems, sink);
    run<LD_CS, true>("cs", g, table, ids, out, nitems, sink);
    run<LD_NC_L2_64, true>("nc.L2::64B", g, table, ids, out, nitems, sink);
    run<LD_RELAXED, true>("relaxed.gpu", g, table, ids, out, nitems, sink);
  }
  return 0;
}
