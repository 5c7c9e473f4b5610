// gather_ceiling.cu — what can a BARE random-row copy reach on this GPU?  (diagnostic, not product)
//
// The fused gather's roofline fraction is quoted against the sequential copy peak, but its reads are
// random 64-byte rows.  This standalone program measures, on tables far larger than L2:
//   seq_copy            float4 grid-stride copy (the "peak" analogue, read + write)
//   row_copy<ROWB,U>    out[i,:] = table[ids[i],:] for ROWB-byte rows, U independent row loads in
//                       flight per lane, int32 ids (no field descriptors, no bounds checks)
//   row_read<64,U>      the same reads, no output stream (pure random-read rate)
// and prints one JSON line per variant.  Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3
//   -o tools/gather_ceiling tools/gather_ceiling.cu      Run:  tools/gather_ceiling [table_GB] [items_M]
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
__global__ void __launch_bounds__(256) seq_copy(const float4* __restrict__ in, float4* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
    out[i] = in[i];
}

// LPR = lanes per row = ROWB / 16
template <int ROWB, int U, bool WRITE>
__global__ void __launch_bounds__(256) row_copy(const float4* __restrict__ table, const int32_t* __restrict__ ids,
                                                float4* __restrict__ out, int64_t nitems, float* sink) {
  constexpr int LPR = ROWB / 16;
  const int sub = threadIdx.x % LPR;
  const int64_t ngroups = ((int64_t) gridDim.x * blockDim.x) / LPR;
  const int64_t group = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  float acc = 0.f;
  for (int64_t base = group; base < nitems; base += ngroups * U) {
    int64_t row[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t it = base + (int64_t) u * ngroups;
      row[u] = it < nitems ? ids[it] : -1;
    }
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      v[u] = row[u] >= 0 ? table[row[u] * LPR + sub] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t it = base + (int64_t) u * ngroups;
      if (WRITE) { if (it < nitems) out[it * LPR + sub] = v[u]; }
      else acc += v[u].x + v[u].w;
    }
  }
  if (!WRITE && acc == 123456.789f) *sink = acc;
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

template <int ROWB, int U, bool WRITE>
static void run(const float4* table, const int32_t* ids, float4* out, int64_t nitems, float* sink, int ctas_per_sm) {
  constexpr int LPR = ROWB / 16;
  const int grid = 148 * ctas_per_sm;
  const float ms = time_ms([&] { row_copy<ROWB, U, WRITE><<<grid, 256>>>(table, ids, out, nitems, sink); }, 5);
  const double bytes = (double) nitems * (4.0 + ROWB + (WRITE ? ROWB : 0));
  printf("{\"kernel\": \"%s\", \"row_bytes\": %d, \"unroll\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n",
         WRITE ? "row_copy" : "row_read", ROWB, U, ctas_per_sm, ms, bytes / ms / 1e6);
  fflush(stdout);
  (void) LPR;
}

int main(int argc, char** argv) {
  const double table_gb = argc > 1 ? atof(argv[1]) : 10.0;
  const int64_t nitems = (int64_t) ((argc > 2 ? atof(argv[2]) : 20.4) * 1e6);   // 524288 x 39 = 20.4 M
  const int64_t table_f4 = (int64_t) (table_gb * 1e9 / 16);
  float4 *table, *out;
  int32_t* ids;
  float* sink;
  CK(cudaMalloc(&table, table_f4 * 16));
  CK(cudaMalloc(&out, nitems * 256));          // room for 256-byte rows
  CK(cudaMalloc(&ids, nitems * 4));
  CK(cudaMalloc(&sink, 4));
  fill_kernel<<<148 * 8, 256>>>(table, table_f4);
  CK(cudaDeviceSynchronize());
  {
    const int64_t n = nitems * 4;                // 64 B per item, as the gather moves
    const float ms = time_ms([&] { seq_copy<<<148 * 8, 256>>>(table, out, n); }, 5);
    printf("{\"kernel\": \"seq_copy\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, (double) n * 32 / ms / 1e6);
  }
  // 64-byte rows (D = 16 fp32): the C2 gather
  ids_kernel<<<148 * 8, 256>>>(ids, nitems, (uint32_t) (table_f4 / 4));
  CK(cudaDeviceSynchronize());
  run<64, 1, true>(table, ids, out, nitems, sink, 8);
  run<64, 2, true>(table, ids, out, nitems, sink, 8);
  run<64, 4, true>(table, ids, out, nitems, sink, 8);
  run<64, 8, true>(table, ids, out, nitems, sink, 8);
  run<64, 16, true>(table, ids, out, nitems, sink, 8);
  run<64, 8, true>(table, ids, out, nitems, sink, 4);
  run<64, 8, true>(table, ids, out, nitems, sink, 16);
  run<64, 4, true>(table, ids, out, nitems, sink, 16);
  run<64, 4, false>(table, ids, out, nitems, sink, 8);
  run<64, 8, false>(table, ids, out, nitems, sink, 8);
  run<64, 16, false>(table, ids, out, nitems, sink, 8);
  // wider rows: how much of the gap is the 64-byte access size?
  ids_kernel<<<148 * 8, 256>>>(ids, nitems, (uint32_t) (table_f4 / 8));
  CK(cudaDeviceSynchronize());
  run<128, 8, true>(table, ids, out, nitems, sink, 8);
  run<128, 8, false>(table, ids, out, nitems, sink, 8);
  ids_kernel<<<148 * 8, 256>>>(ids, nitems, (uint32_t) (table_f4 / 16));
  CK(cudaDeviceSynchronize());
  run<256, 8, true>(table, ids, out, nitems, sink, 8);
  run<256, 8, false>(table, ids, out, nitems, sink, 8);
  return 0;
}
