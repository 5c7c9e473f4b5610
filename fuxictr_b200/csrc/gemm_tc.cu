// gemm_tc.cu — TMA-fed tcgen05 (5th-gen tensor core) GEMM with TMEM accumulator and a
// fused epilogue, sm_100a.  The dense contraction behind MLP_Block
// (fuxictr/pytorch/layers/blocks/mlp_block.py:74-85), CrossNetV2's d x d Linear
// (fuxictr/pytorch/layers/interactions/cross_net.py:126-129) and CIN's 1x1 Conv1d
// (fuxictr/pytorch/layers/interactions/compressed_interaction_net.py:72).
//
//   C[m, n] = epi( sum_k A(m, k) * B(n, k) )
// Each operand is either K-major (memory (rows, K), K contiguous) or MN-major (memory (K, rows),
// rows contiguous): the tensor core reads both through its shared-memory matrix descriptor, so
// the dgrad (dX = dZ W) and wgrad (dW = dZ^T X) contractions of a Linear layer consume W, dZ and X
// exactly as they lie in memory — no transpose pass.
//
// Arithmetic: tcgen05.mma kind::tf32 reads the fp32 operands straight from shared memory
// (the tensor core ignores the low 13 mantissa bits) and accumulates in fp32 in TMEM.
//   precision 1 ("tf32"):   one pass, ~1e-3 relative — at least the bf16 the config names.
//   precision 3 ("tf32x3"): error-compensated 3xTF32: with x = big(x) + small(x),
//        A.B ~= A_big.B_big + A_big.B_small + A_small.B_big
//     where big() is the hardware truncation itself and small = rna_tf32(x - big(x)) — fp32-class accuracy
//     (the 1e-5 parity bar) on tensor cores.  The small tiles are made in shared memory by the helper
//     warps (B2_GEMM_X3_INLINE: the GEMM then reads only the fp32 operands) or loaded from HBM (a_small /
//     b_small of the descriptor, produced by b2_split_tf32 or a previous epilogue).
//   bf16 operands (elem_dtype B2_BF16): kind::f16, 64-element k-blocks, fp32 accumulation.
//
// Structure (persistent: at most one CTA per SM, striding over the 128 x BN output tiles x K splits):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled tiles into a 2-4-stage mbarrier ring
//   warp 1      TMEM allocator + tcgen05.mma issuer (one elect.sync lane), tcgen05.commit -> mbarriers
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> smem transpose -> fused epilogue ->
//               TMA tensor store of the finished 32 x 32 chunk (register stores for split-K / accumulate / c_pre)
//   warps 6-9   helpers: (inline 3xTF32) A -> As, B -> Bs in shared memory, then every second epilogue chunk
//   TMEM holds 1 or 2 accumulator stages (full/empty mbarriers), so a CTA with several tiles never
//   drains its pipelines between them; in one-tile CTAs the epilogue's patches reuse the dead ring.
//   Kernels of a step are chained with programmatic dependent launch (prologue under the predecessor's tail).
// Where the time goes (probe build): SS-mode operand fetch from shared memory and the
// short MMA issue queue bound the main loop; the output burst bounds the epilogue.
// Every mbarrier wait is bounded: a pipeline bug traps with a message instead of hanging the GPU.
#include "b2_common.cuh"
#include <string.h>
#include <cstdlib>
#include <cuda.h>  // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time

namespace tc {
constexpr int BM = 128;          // UMMA M (cta_group::1): TMEM lane == output row
constexpr int BK = 32;           // fp32 elements per k-block == one 128-byte swizzle row (bf16: 64)
constexpr int UMMA_K_BYTES = 32; // kind::tf32: K = 8 elements of 4 bytes per instruction
constexpr int MAX_STAGES = 4;  // 4 stages of (A,B) for one pass; 3 stages of (A,As,B,Bs) for 3xTF32
constexpr int A_BYTES = BM * 128;
constexpr int NTHREADS = 320;   // producer, MMA issuer, 4 epilogue warps, 4 helper warps (second half of the epilogue
                                // columns; with B2_GEMM_X3_INLINE they first make the 3xTF32 small parts in shared memory)
constexpr int PATCH_BYTES = 8 * 32 * 33 * 4;   // epilogue transpose patches (one per epilogue / helper warp)

// Timing probes (tools/gemm_probe.py, tools/gemm_trace.py): compiled in only with -DB2_GEMM_PROBE
// (B2_BUILD_PROBE=1 python -m fuxictr_b200.build); the product build has no debug hooks.
#ifdef B2_GEMM_PROBE
#define B2_DBG(bit) ((p.dbg & (bit)) != 0)
#define B2_STAMP(cond, idx) do { if (p.trace != nullptr && blockIdx.x == 0 && (cond)) p.trace[idx] = clock64(); } while (0)
#else
#define B2_DBG(bit) false
#define B2_STAMP(cond, idx) do { } while (0)
#endif

struct Params {
  CUtensorMap map_a[2];   // [0] the operand, [1] its 3xTF32 small part
  CUtensorMap map_b[2];
  CUtensorMap map_c;      // C as a (N, M) tensor with 32 x 32 boxes: the epilogue's TMA store (tma_store != 0)
  float* c;
  float* c_small;         // optional: tf32_small(C) for the consumer's 3xTF32 operand
  float* c_pre;           // optional: acc + bias BEFORE mul/add/act (CrossNetV2 saves it for its backward)
  int64_t ldc;
  const float* bias;
  const float* mul;
  const float* add;
  const float* ybwd;      // optional (M, N) ld = ldc: C = act_bwd'(ybwd) * (...)   (activation backward)
  float* colsum;          // optional (N): += column sums of C (bias gradient)
  int M, N, K, bn, nseg, act, act_bwd, beta, kb_per_split;
  int a_mn, b_mn;         // operand is MN-major (memory (K, rows))
  int esz;                // operand element bytes: 4 = fp32 read as tf32 (kind::tf32), 2 = bf16 (kind::f16)
  int64_t ld_aux;         // leading dimension of c_small (fp32 small part, or the bf16 copy of C when esz == 2)
  int nmain;      // TMEM accumulators for the main (big x big) product: its K range is cut in nmain chunks
  int tiles_m, tiles_n, splits;   // tile grid; CTAs stride over tiles_m * tiles_n * splits work items
  int nacc;       // accumulator stages in TMEM (2: the epilogue of a tile overlaps the next main loop)
  int stages;     // operand ring depth (<= MAX_STAGES), chosen by the host to fit 227 KB
  long long* trace;   // probe build only: CTA 0 writes clock64() stamps of its loops here (B2_GEMM_TRACE)
  int dbg;        // probe build only: B2_GEMM_DBG bits (timing experiments; results are wrong when set)
  int tma_store;  // the epilogue hands finished 32 x 32 chunks to cp.async.bulk.tensor stores (plain C output only)
  int inline_split;  // 3xTF32 with the small parts computed in shared memory by warps 6..9 (no As/Bs in HBM)
  int tmem_cols;  // power of two >= (nmain + (nseg > 1)) * bn
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: ~2 s at 2 GHz, then report and trap (never hang the device).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int which) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("b2 tc_gemm: mbarrier wait timed out (role %d, block %d,%d,%d)\n", which, blockIdx.x,
             blockIdx.y, blockIdx.z);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// One elected lane of a fully converged warp (elect.sync): unlike `lane == 0`, the compiler knows the
// enclosing control flow is warp-uniform, so descriptors / coordinates stay in uniform registers and
// every UTCHMMA / UTMALDG is a single instruction instead of a per-instruction election loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// K-major, 128-byte swizzle, tile rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t) ((addr & 0x3FFFFu) >> 4);  // start address  [0,14)
  d |= (uint64_t) 0 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t) (1024 >> 4) << 32;         // stride byte offset [32,46)
  d |= (uint64_t) 1 << 46;                   // descriptor version (sm_100)
  d |= (uint64_t) 2 << 61;                   // layout: SWIZZLE_128B
  return d;
}
// MN-major: the tile is a row of [128 bytes of MN-elements x bke k-rows] boxes (what one TMA box writes:
// k-row r at r * 128 B), boxes box_bytes apart = LBO.
//   16-bit operands: SWIZZLE_128B, swizzle atom = 8 k-rows (1024 B) = SBO.
//   32-bit operands (tf32): the ONLY MN-major layout the tensor core accepts is SWIZZLE_128B_BASE32B
//   (32-byte swizzle granules, atom = 4 k-rows = 512 B = SBO); its TMA twin is SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr, uint32_t box_bytes, int esz) {
  uint64_t d = 0;
  d |= (uint64_t) ((addr & 0x3FFFFu) >> 4);      // start address  [0,14)
  d |= (uint64_t) (box_bytes >> 4) << 16;        // leading byte offset [16,30): next 128-byte-wide MN block
  d |= (uint64_t) ((esz == 4 ? 512 : 1024) >> 4) << 32;   // stride byte offset [32,46): next swizzle atom along K
  d |= (uint64_t) 1 << 46;                       // descriptor version (sm_100)
  d |= (uint64_t) (esz == 4 ? 1 : 2) << 61;      // layout: SWIZZLE_128B_BASE32B (1) / SWIZZLE_128B (2)
  return d;
}
__device__ __forceinline__ float tf32_small(float v) { return b2_tf32_small(v); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One lane's share of the epilogue: column n of 32 consecutive rows (mrow0 ...), values in t[].
// `store_c` false: everything but the store of C itself (the caller hands the chunk to a TMA store).
__device__ __forceinline__ void epilogue_store(const Params& p, float (&t)[32], int mrow0, int n, bool split, bool store_c) {
  float* cp = p.c + (int64_t) mrow0 * p.ldc + n;
  if (split) {
#pragma unroll
    for (int r = 0; r < 32; ++r)
      if (mrow0 + r < p.M) b2_red_add(cp + (int64_t) r * p.ldc, t[r]);
  } else {
    if (p.c_pre != nullptr) {
      float* pp = p.c_pre + (int64_t) mrow0 * p.ldc + n;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (mrow0 + r < p.M) pp[(int64_t) r * p.ldc] = t[r];
    }
    if (p.mul != nullptr) {
      const float* mp = p.mul + (int64_t) mrow0 * p.ldc + n;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (mrow0 + r < p.M) t[r] *= __ldg(mp + (int64_t) r * p.ldc);
    }
    if (p.add != nullptr) {
      const float* ap = p.add + (int64_t) mrow0 * p.ldc + n;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (mrow0 + r < p.M) t[r] += __ldg(ap + (int64_t) r * p.ldc);
    }
    if (p.act == B2_ACT_RELU) {
#pragma unroll
      for (int r = 0; r < 32; ++r) t[r] = fmaxf(t[r], 0.f);
    } else if (p.act == B2_ACT_SIGMOID) {
#pragma unroll
      for (int r = 0; r < 32; ++r) t[r] = 1.f / (1.f + expf(-t[r]));
    }
    if (p.ybwd != nullptr) {   // activation backward of the PRODUCER of this gradient, fused
      const float* yp = p.ybwd + (int64_t) mrow0 * p.ldc + n;
      if (p.act_bwd == B2_ACT_RELU) {
#pragma unroll
        for (int r = 0; r < 32; ++r)
          if (mrow0 + r < p.M) t[r] = (__ldg(yp + (int64_t) r * p.ldc) > 0.f) ? t[r] : 0.f;
      } else if (p.act_bwd == B2_ACT_SIGMOID) {
#pragma unroll
        for (int r = 0; r < 32; ++r)
          if (mrow0 + r < p.M) {
            const float yv = __ldg(yp + (int64_t) r * p.ldc);
            t[r] = t[r] * ((1.f - yv) * yv);
          }
      }
    }
    if (p.beta) {
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (mrow0 + r < p.M) t[r] += cp[(int64_t) r * p.ldc];
    }
    if (store_c) {
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (mrow0 + r < p.M) cp[(int64_t) r * p.ldc] = t[r];   // 128 contiguous bytes per row
    }
    if (p.c_small != nullptr && p.esz == 4) {   // the consumer's 3xTF32 small part, produced where C is produced
      float* sp = p.c_small + (int64_t) mrow0 * p.ld_aux + n;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (mrow0 + r < p.M) sp[(int64_t) r * p.ld_aux] = tf32_small(t[r]);
    } else if (p.c_small != nullptr) {          // bf16 mode: the consumer's bf16 operand (round-to-nearest-even)
      __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(p.c_small) + (int64_t) mrow0 * p.ld_aux + n;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (mrow0 + r < p.M) sp[(int64_t) r * p.ld_aux] = __float2bfloat16_rn(t[r]);
    }
    if (p.colsum != nullptr) {    // bias gradient: this lane owns column n of 32 rows
      float cs = 0.f;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (mrow0 + r < p.M) cs += t[r];
      b2_red_add(p.colsum + n, cs);
    }
  }
}

// Persistent tile loop: CTA c works on tiles c, c + gridDim.x, ... (one tile per CTA when the problem
// has no more tiles than SMs).  Three pipelines run through all of a CTA's tiles without draining:
//   smem ring      full[s] / empty[s]            TMA producer  <->  MMA issuer
//   TMEM stages    tmem_full[a] / tmem_empty[a]  MMA issuer    <->  epilogue warps (a < nacc = 1 or 2)
// so the barrier / TMEM / descriptor prologue is paid once per CTA, the producer prefetches the next
// tile's operands during an epilogue, and with nacc = 2 the epilogue of tile j overlaps the main loop
// of tile j + 1.
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tf32_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte aligned bases: align by hand (1 KB of slack is requested).
  // (pointer arithmetic on smem_raw keeps the shared address space visible to the compiler)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t b_bytes = (uint32_t) p.bn * 128u;
  const bool x3 = p.nseg > 1;
  const int STAGES = p.stages;
  const uint32_t stage_bytes = (x3 ? 2u : 1u) * (A_BYTES + b_bytes);
  const uint32_t off_as = A_BYTES, off_b = (x3 ? 2u : 1u) * A_BYTES, off_bs = off_b + b_bytes;
  // The 8 epilogue transpose patches (32 x 33 floats each) follow the ring — or, when every CTA has exactly
  // one tile, lie ON the ring: the accumulator is complete only after every MMA has read its operands and
  // nothing is loaded afterwards, so the ring is dead by then and its bytes buy one more stage instead.
  const bool single_tile = p.tiles_m * p.tiles_n * p.splits <= (int) gridDim.x;
  float* patch_base = reinterpret_cast<float*>(smem + (single_tile ? 0u : STAGES * stage_bytes));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes + (single_tile ? 0 : PATCH_BYTES));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + MAX_STAGES),
                 tfull0 = smem_u32(bars + 2 * MAX_STAGES), tempty0 = smem_u32(bars + 2 * MAX_STAGES + 2),
                 conv0 = smem_u32(bars + 2 * MAX_STAGES + 4);
  const bool inl = p.inline_split != 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bke = 128 / p.esz;                      // k elements per k-block: one 128-byte swizzle row
  const int mnb = 128 / p.esz;                      // MN elements per MN-major box (128 bytes wide)
  const uint32_t box_bytes = (uint32_t) bke * 128u; // one MN-major box: bke k-rows x 128 B
  const int num_kb_total = (p.K + bke - 1) / bke;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int total_tiles = tiles_mn * p.splits;
  const int acc_cols = (p.nmain + (x3 ? 1 : 0)) * p.bn;   // TMEM columns of one accumulator stage
  // Warps 6..9 take every second 32-column chunk of the epilogue (TMEM lane quadrant = warp % 4, like warps
  // 2..5).  With the inline split they are the converters first, so they help only when this CTA has a
  // single tile (otherwise they are already converting the next tile's operands).
  const bool helpers = (!inl || total_tiles <= (int) gridDim.x) && !B2_DBG(32);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < ((x3 && !inl) ? 2 : 1); ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_a[s])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_b[s])) : "memory");
    }
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.map_c)) : "memory");
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
      mbar_init(conv0 + 8 * s, 4);      // one arrival per converter warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, helpers ? 8 : 4);    // one arrival per warp that reads the accumulator
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // whole warp: allocate TMEM columns for the accumulator stage(s)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t) p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  B2_STAMP(threadIdx.x == 0, 640);
  // PDL: everything above touched only shared memory, TMEM and the kernel parameters; from here on the
  // predecessor's outputs are read.  Let the successor begin ITS prologue once every CTA got this far.
  b2_pdl_trigger();
  b2_pdl_wait();

  // One warp's share of a tile's epilogue: 32-column chunks first, first + step, ... of TMEM lane quadrant
  // warp % 4.  Each warp owns a 32 x 33-float patch to transpose its TMEM rows, so that one store
  // instruction writes 128 contiguous bytes of ONE output row instead of 16 bytes of 32 different rows
  // (partial-sector writes to untouched lines cost an L2 fill each — measured 19 us per tile before this).
  auto epilogue_tile = [&](int t, int j, int first, int step) {
    const int q = warp & 3;
    const bool split = p.splits > 1;
    float* patch = patch_base + (warp - 2) * (32 * 33);
    const int z = t / tiles_mn, rr = t - z * tiles_mn;
    const int m0 = (rr / p.tiles_n) * BM, n0 = (rr % p.tiles_n) * p.bn;
    const int kb_begin = z * p.kb_per_split;
    const int nkb = min(num_kb_total, kb_begin + p.kb_per_split) - kb_begin;
    const int nmain = min(p.nmain, nkb);
    const int nslots = nmain + (x3 ? 1 : 0);
    const int acc = (p.nacc == 2) ? (j & 1) : 0;
    const uint32_t use = (uint32_t) (p.nacc == 2 ? (j >> 1) : j);
    B2_STAMP(lane == 0, 512 + 4 * warp + 0);
    mbar_wait(tfull0 + 8 * acc, use & 1u, 2);
    B2_STAMP(lane == 0, 512 + 4 * warp + 1);
    tc_fence_after();
    const uint32_t tacc = tmem_base + (uint32_t) (acc * acc_cols) + ((uint32_t) (q * 32) << 16);
    if (first >= p.bn) {          // a helper warp with no chunk in a narrow tile still owes its arrival
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
    }
    for (int c0 = first; c0 < p.bn; c0 += step) {
      uint32_t v[32];
      __syncwarp();  // tcgen05.ld is warp-collective; also fences the previous patch reads
      tmem_ld32(tacc + (uint32_t) c0, v);
      for (int sl = 1; sl < nslots; ++sl) {  // fp32 round-to-nearest sum of the accumulation chains
        uint32_t w[32];
        tmem_ld32(tacc + (uint32_t) (sl * p.bn + c0), w);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = __float_as_uint(__uint_as_float(v[jj]) + __uint_as_float(w[jj]));
      }
      if (c0 + step >= p.bn) {      // this warp's last TMEM read of the tile: hand the accumulator stage back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      }
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) patch[lane * 33 + jj] = __uint_as_float(v[jj]);  // row = lane
      __syncwarp();
      const int n = n0 + c0 + lane;  // this lane's output column for the whole chunk
      const bool n_ok = n < p.N;
      const float bv = (n_ok && p.bias != nullptr && z == 0) ? __ldg(p.bias + n) : 0.f;
      float tt[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) tt[r] = patch[r * 33 + lane] + bv;  // 32 independent LDS in flight
      const int mrow0 = m0 + q * 32;
      if (!p.tma_store) {
        if (n_ok && !B2_DBG(2)) epilogue_store(p, tt, mrow0, n, split, true);
      } else {
        // TMA store: the finished chunk goes back into the warp's patch as a dense 32 x 32 box (row = output
        // row, 128 B per row) and ONE bulk tensor store writes it; rows >= M and columns >= N are clipped by
        // the tensor map, so nothing here is predicated on the tile edges.
        if (n_ok) epilogue_store(p, tt, mrow0, n, split, false);
        __syncwarp();                  // every lane has read its column out of the 33-pitch patch
#pragma unroll
        for (int r = 0; r < 32; ++r) patch[r * 32 + lane] = tt[r];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && !B2_DBG(2)) {
          tma_store_2d(&p.map_c, smem_u32(patch), n0 + c0, mrow0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the patch is rewritten by the next chunk
        }
        __syncwarp();
      }
    }
    if (p.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    B2_STAMP(lane == 0, 512 + 4 * warp + 2);
  };

  if (warp == 0) {
    // ---------------- TMA producer (the warp loops converged; one elected lane issues) ----------------
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int z = t / tiles_mn, r = t - z * tiles_mn;
      const int m0 = (r / p.tiles_n) * BM, n0 = (r % p.tiles_n) * p.bn;
      const int kb_begin = z * p.kb_per_split;
      const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        B2_STAMP(lane == 0 && kb - kb_begin < 60, 4 * (kb - kb_begin) + 0);
        mbar_wait(empty0 + 8 * stage, phase ^ 1, 0);
        B2_STAMP(lane == 0 && kb - kb_begin < 60, 4 * (kb - kb_begin) + 1);
        const uint32_t a_dst = smem_base + stage * stage_bytes;
        const uint32_t full = full0 + 8 * stage;
        if (B2_DBG(16) && elect_one()) mbar_arrive(full);      // probe: no operand traffic at all
        if (!B2_DBG(16) && elect_one()) {
          mbar_expect_tx(full, inl ? (A_BYTES + b_bytes) : stage_bytes);
          // every operand tile is loaded ONCE per k-block; 3xTF32 reuses them for its 3 products.
          // K-major operand: one box (128 B of k x rows).  MN-major operand: one box per 128 B of rows
          // (coordinates {row, k}); out-of-range boxes arrive zero-filled.
          for (int s = 0; s < ((x3 && !inl) ? 2 : 1); ++s) {
            const uint32_t a_t = a_dst + (s ? off_as : 0u), b_t = a_dst + (s ? off_bs : off_b);
            if (!p.a_mn) {
              tma_load_2d(a_t, &p.map_a[s], full, kb * bke, m0);
            } else {
              for (int j = 0; j < BM / mnb; ++j) tma_load_2d(a_t + j * box_bytes, &p.map_a[s], full, m0 + mnb * j, kb * bke);
            }
            if (!p.b_mn) {
              tma_load_2d(b_t, &p.map_b[s], full, kb * bke, n0);
            } else {
              for (int j = 0; j < p.bn / mnb; ++j) tma_load_2d(b_t + j * box_bytes, &p.map_b[s], full, n0 + mnb * j, kb * bke);
            }
          }
        }
        __syncwarp();
        B2_STAMP(lane == 0 && kb - kb_begin < 60, 4 * (kb - kb_begin) + 2);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (the warp loops converged; one elected lane issues) ----------------
    // instruction descriptor: D=f32, A=B=tf32 | bf16, operand majors, N = bn, M = 128
    const uint32_t fmt = (p.esz == 2) ? 1u : 2u;   // operand format: kind::f16 1 = BF16; kind::tf32 2 = TF32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t) (p.a_mn ? 1 : 0) << 15) |
                           ((uint32_t) (p.b_mn ? 1 : 0) << 16) | ((uint32_t) (p.bn >> 3) << 17) |
                           ((uint32_t) (BM >> 4) << 24);
    // one instruction consumes 32 bytes of K per row (8 tf32 / 16 bf16 elements): 32 bytes along a
    // K-major swizzle row (+2 in the (addr >> 4) field), or 8 / 16 k-rows of an MN-major tile
    const uint32_t mn_step = (p.esz == 2) ? 2048u : 1024u;
    const uint32_t ak = p.a_mn ? (mn_step >> 4) : (uint32_t) (UMMA_K_BYTES >> 4);
    const uint32_t bk = p.b_mn ? (mn_step >> 4) : (uint32_t) (UMMA_K_BYTES >> 4);
    // The tensor core adds each instruction's 8 products into the fp32 accumulator with
    // truncation, so rounding error grows with the length of one accumulation chain and with
    // the magnitude of the accumulator.  For 3xTF32 the K range of the main product is
    // therefore cut into `nmain` chains held in separate TMEM column ranges, and the two small
    // correction products get a range of their own; the epilogue adds the ranges in fp32 RN.
    //
    // The elected lane is the instruction stream behind every MMA of the CTA, so its scalar work per
    // k-block is on the critical path: descriptors are built once (stage 0) and advanced by 32-bit
    // adds on their low word (the 14-bit start-address field never carries), the chain boundaries
    // are tracked incrementally, and the bf16 / TF32 / 3xTF32 loops are separate.
    const uint64_t a0 = p.a_mn ? make_smem_desc_mn(smem_base, box_bytes, p.esz) : make_smem_desc(smem_base);
    const uint64_t b0 = p.b_mn ? make_smem_desc_mn(smem_base + off_b, box_bytes, p.esz) : make_smem_desc(smem_base + off_b);
    const uint64_t as0 = p.a_mn ? make_smem_desc_mn(smem_base + off_as, box_bytes, p.esz) : make_smem_desc(smem_base + off_as);
    const uint64_t bs0 = p.b_mn ? make_smem_desc_mn(smem_base + off_bs, box_bytes, p.esz) : make_smem_desc(smem_base + off_bs);
    const uint32_t a_hi = (uint32_t) (a0 >> 32), b_hi = (uint32_t) (b0 >> 32);
    const uint32_t as_hi = (uint32_t) (as0 >> 32), bs_hi = (uint32_t) (bs0 >> 32);
    const uint32_t stage_units = stage_bytes >> 4;
#define B2_DESC(hi, lo) ((((uint64_t) (hi)) << 32) | (uint64_t) (uint32_t) (lo))
    int stage = 0;
    uint32_t phase = 0, so = 0;      // so: this stage's offset in descriptor units
    int j = 0;                        // this CTA's tile counter
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      const int z = t / tiles_mn;
      const int kb_begin = z * p.kb_per_split;
      const int nkb = min(num_kb_total, kb_begin + p.kb_per_split) - kb_begin;
      const int nmain = min(p.nmain, nkb);        // a short last K split may hold fewer k-blocks than chains
      const int acc = (p.nacc == 2) ? (j & 1) : 0;
      const uint32_t use = (uint32_t) (p.nacc == 2 ? (j >> 1) : j);
      mbar_wait(tempty0 + 8 * acc, (use & 1u) ^ 1u, 3);    // the epilogue drained this accumulator stage
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t) (acc * acc_cols);
      const uint32_t d_corr = tacc + (uint32_t) (nmain * p.bn);
      // chain c covers k-blocks [ceil(c nkb / nmain), ceil((c+1) nkb / nmain)): divisions only at the boundaries
      int slot = 0, this_start = 0, next_start = (nkb + nmain - 1) / nmain;
      const uint32_t ready0 = inl ? conv0 : full0;           // operands (and their small parts) are in place
      mbar_wait(ready0 + 8 * stage, phase, 1);
      for (int i = 0; i < nkb; ++i) {
        if (i == next_start) {
          ++slot;
          this_start = next_start;
          next_start = ((slot + 1) * nkb + nmain - 1) / nmain;
        }
        const uint32_t keep = (i == this_start) ? 0u : 1u;      // first k-block of a chain overwrites
        const uint32_t d_main = tacc + (uint32_t) (slot * p.bn);
        const uint32_t al = (uint32_t) a0 + so, bl = (uint32_t) b0 + so;
        const uint32_t asl = (uint32_t) as0 + so, bsl = (uint32_t) bs0 + so;
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
        // The tensor pipe's instruction queue is short: whatever this thread does between the last MMA of
        // one k-block and the first of the next is a bubble (measured ~380 cycles per k-block against 576 of
        // MMA work).  So each k-block is issued in two halves and the look at the NEXT stage's barrier sits
        // between them, under the queued instructions of the first half.
#define B2_ISSUE_HALF(K0)                                                                                      \
        if (elect_one()) {                                                                                     \
          if (B2_DBG(4)) {        /* probe: no tensor work, only the pipeline hand-offs */                      \
          } else if (p.esz == 2) {                                                                             \
            _Pragma("unroll") for (int k = (K0); k < (K0) + 2; ++k)                                            \
              umma_bf16(d_main, B2_DESC(a_hi, al + k * ak), B2_DESC(b_hi, bl + k * bk), idesc, (k > 0) ? 1u : keep); \
          } else if (!x3) {                                                                                    \
            _Pragma("unroll") for (int k = (K0); k < (K0) + 2; ++k)                                            \
              umma_tf32(d_main, B2_DESC(a_hi, al + k * ak), B2_DESC(b_hi, bl + k * bk), idesc, (k > 0) ? 1u : keep); \
          } else {                                                                                             \
            _Pragma("unroll") for (int k = (K0); k < (K0) + 2; ++k) {                                          \
              const uint64_t ad = B2_DESC(a_hi, al + k * ak), bd = B2_DESC(b_hi, bl + k * bk);                 \
              umma_tf32(d_main, ad, bd, idesc, (k > 0) ? 1u : keep);                                   /* A_big . B_big */   \
              umma_tf32(d_corr, ad, B2_DESC(bs_hi, bsl + k * bk), idesc, (i > 0 || k > 0) ? 1u : 0u); /* A_big . B_small */ \
              umma_tf32(d_corr, B2_DESC(as_hi, asl + k * ak), bd, idesc, 1u);                          /* A_small . B_big */ \
            }                                                                                                  \
          }                                                                                                    \
          if ((K0) == 2) {                                                                                     \
            umma_commit(empty0 + 8 * stage);                      /* frees this smem slot once the MMAs have read it */ \
            if (i == nkb - 1) umma_commit(tfull0 + 8 * acc);      /* ... and the last one: accumulator complete */      \
          }                                                                                                    \
        }                                                                                                      \
        __syncwarp();
        B2_STAMP(lane == 0 && i < 60, 256 + 4 * i + 0);
        B2_ISSUE_HALF(0)
        B2_STAMP(lane == 0 && i < 60, 256 + 4 * i + 1);
        // non-blocking look at the next stage under the queued first half; block for it only after the
        // second half is queued too (a blocking wait here would hold back MMAs whose operands are present)
        const bool more = i + 1 < nkb;
        const bool next_ready = more && __all_sync(0xffffffffu, mbar_try_wait(ready0 + 8 * nstage, nphase));
        B2_ISSUE_HALF(2)
#undef B2_ISSUE_HALF
        B2_STAMP(lane == 0 && i < 60, 256 + 4 * i + 2);
        if (more && !next_ready) mbar_wait(ready0 + 8 * nstage, nphase, 1);
        B2_STAMP(lane == 0 && i < 60, 256 + 4 * i + 3);
        so += stage_units;
        stage = nstage;
        phase = nphase;
        if (stage == 0) so = 0;
      }
    }
#undef B2_DESC
  } else if (warp < 6) {
    // ---------------- epilogue warps 2..5: TMEM lane quadrant = warp % 4 ----------------
    int j = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j)
      epilogue_tile(t, j, 0, helpers ? 64 : 32);
  } else {
    // ---------------- helper warps 6..9 ----------------
    // Inline 3xTF32: small = x - tf32(x), element for element.  The small tiles have their operand's own
    // shared-memory layout (whatever the swizzle), so the conversion is a flat 16-byte-per-lane pass:
    // A -> As, B -> Bs.  Generic-proxy stores are made visible to the tensor core (async proxy) by
    // fence.proxy.async before the arrival the MMA warp waits on.  Then (or only) the odd epilogue chunks.
    const int cw = threadIdx.x - 6 * 32;          // 0..127
    const int b_chunks = (int) (b_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int j = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      const int z = t / tiles_mn;
      const int kb_begin = z * p.kb_per_split;
      const int nkb = min(num_kb_total, kb_begin + p.kb_per_split) - kb_begin;
      for (int i = 0; inl && i < nkb; ++i) {
        mbar_wait(full0 + 8 * stage, phase, 4);
        uint8_t* sb = smem + (size_t) stage * stage_bytes;
        if (!B2_DBG(1)) {
          float4 v[A_BYTES / 16 / 128];
#pragma unroll
          for (int u = 0; u < A_BYTES / 16 / 128; ++u) v[u] = *reinterpret_cast<const float4*>(sb + 16 * (cw + 128 * u));
#pragma unroll
          for (int u = 0; u < A_BYTES / 16 / 128; ++u) {
            float4 w;
            w.x = tf32_small(v[u].x); w.y = tf32_small(v[u].y); w.z = tf32_small(v[u].z); w.w = tf32_small(v[u].w);
            *reinterpret_cast<float4*>(sb + off_as + 16 * (cw + 128 * u)) = w;
          }
        }
        for (int c = cw; c < (B2_DBG(1) ? 0 : b_chunks); c += 256) {
          const float4 v0 = *reinterpret_cast<const float4*>(sb + off_b + 16 * c);
          const bool two = c + 128 < b_chunks;
          float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (two) v1 = *reinterpret_cast<const float4*>(sb + off_b + 16 * (c + 128));
          float4 w;
          w.x = tf32_small(v0.x); w.y = tf32_small(v0.y); w.z = tf32_small(v0.z); w.w = tf32_small(v0.w);
          *reinterpret_cast<float4*>(sb + off_bs + 16 * c) = w;
          if (two) {
            w.x = tf32_small(v1.x); w.y = tf32_small(v1.y); w.z = tf32_small(v1.z); w.w = tf32_small(v1.w);
            *reinterpret_cast<float4*>(sb + off_bs + 16 * (c + 128)) = w;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(conv0 + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (helpers) epilogue_tile(t, j, 32, 64);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t) p.tmem_cols)
                 : "memory");
  }
}

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ small, int64_t n) {
  b2_pdl_wait();
  b2_pdl_trigger();
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t) gridDim.x * blockDim.x) {
    small[i] = tf32_small(x[i]);
  }
}

// out (cols, rows) = in (rows, cols)^T through a padded shared tile; optionally also the
// small() part of the transposed values (for the 3xTF32 operands of dgrad / wgrad).
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, int64_t rows, int64_t cols, int64_t ld_in,
                 float* __restrict__ out, int64_t ld_out, float* __restrict__ out_small) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t c0 = (int64_t) blockIdx.x * 32, r0 = (int64_t) blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < rows && c < cols) ? __ldg(in + r * ld_in + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t c = c0 + ty + i, r = r0 + tx;  // out[c, r]
    if (c < cols && r < rows) {
      const float v = tile[tx][ty + i];
      out[c * ld_out + r] = v;
      if (out_small != nullptr)
        out_small[c * ld_out + r] = tf32_small(v);
    }
  }
}

// ---------------------------------------------------------------------------------
// Operand preparation for the K-major tensor-core GEMMs, one pass over a (R, C) matrix:
//   v      = act'(y) * x          (act-backward fused when y != NULL; y is the activation OUTPUT)
//   out    = v            (R, C)            out_small  = tf32_small(v)
//   outT   = v^T          (C, R)            outT_small = tf32_small(v^T)
//   colsum[c] += sum_r v[r, c]              (bias gradient)
// Any output pointer may be NULL.  32 x 32 tiles through padded shared memory.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_operand_kernel(const float* __restrict__ x, const float* __restrict__ y, int act, int64_t R,
                    int64_t C, int64_t ld_in, float* __restrict__ out, float* __restrict__ out_small,
                    float* __restrict__ outT, float* __restrict__ outT_small, float* __restrict__ colsum) {
  b2_pdl_wait();
  b2_pdl_trigger();
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t c0 = (int64_t) blockIdx.x * 32, r0 = (int64_t) blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      v = __ldg(x + r * ld_in + c);
      if (y != nullptr) {
        const float yv = __ldg(y + r * ld_in + c);
        if (act == B2_ACT_RELU) v = (yv > 0.f) ? v : 0.f;
        else if (act == B2_ACT_SIGMOID) v = v * ((1.f - yv) * yv);
        else if (act == B2_PREP_MUL) v = v * yv;
      }
      if (out != nullptr) out[r * C + c] = v;
      if (out_small != nullptr) out_small[r * C + c] = tf32_small(v);
    }
    tile[ty + i][tx] = v;
  }
  __syncthreads();
  if (outT != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      const int64_t c = c0 + ty + i, r = r0 + tx;  // outT[c, r]
      if (c < C && r < R) {
        const float v = tile[tx][ty + i];
        outT[c * R + r] = v;
        if (outT_small != nullptr) outT_small[c * R + r] = tf32_small(v);
      }
    }
  }
  if (colsum != nullptr && ty == 0) {
    const int64_t c = c0 + tx;
    if (c < C) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) t += tile[i][tx];
      b2_red_add(colsum + c, t);
    }
  }
}

// ---------------------------------------------------------------------------------
// The N = 1 output head of an MLP (Linear(K, 1)): TMA cannot address a 4-byte row and a 128-wide
// MMA tile would be 1/128 full, so it is a warp-per-row GEMV forward and one fused backward:
//   fwd: y[m] = act(<x[m,:], w> + b)
//   bwd: gz = act'(y) * gy;  gx[m,:] = gz[m] * w;  gw += sum_m gz[m] * x[m,:];  gb += sum_m gz[m]
//   With prev_act != NONE the head's input x IS the previous layer's activation output, so that
//   layer's activation backward is fused here: gx <- prev_act'(x) * gx (= dZ of the previous layer),
//   together with its 3xTF32 small part (gx_small) and its bias gradient gb_prev[k] = sum_m gx[m,k].
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                int64_t M, int K, int act, float* __restrict__ y) {
  b2_pdl_wait();
  b2_pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
  const float bv = (b != nullptr) ? __ldg(b) : 0.f;
  for (int64_t m = warp; m < M; m += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(x + m * K + k), __ldg(w + k), acc);
    acc = b2_warp_sum(acc);
    if (lane == 0) {
      float v = acc + bv;
      if (act == B2_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == B2_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      y[m] = v;
    }
  }
}

// CTA = 256 threads = 8 warps; each CTA owns a contiguous block of rows; lane k-strided columns.
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ y,
                const float* __restrict__ gy, int64_t M, int K, int act, int64_t rows_per_cta,
                float* __restrict__ gx, float* __restrict__ gw, float* __restrict__ gb, int prev_act,
                float* __restrict__ gx_small, float* __restrict__ gb_prev) {
  b2_pdl_wait();
  b2_pdl_trigger();
  extern __shared__ float sgw[];  // K partial sums of gw, then K partial sums of gb_prev
  __shared__ float red[32];
  float* sgp = sgw + K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) { sgw[k] = 0.f; sgp[k] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t) blockIdx.x * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  float gb_acc = 0.f;
  // each warp keeps per-lane partials of gw for its k-strided columns over its rows
  for (int kb = 0; kb < K; kb += 32 * 8) {       // 8 columns per lane per pass
    float part[8], cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { part[j] = 0.f; cs[j] = 0.f; }
    for (int64_t m = r0 + warp; m < r1; m += 8) {
      float gz = __ldg(gy + m);
      if (y != nullptr) {
        const float yv = __ldg(y + m);
        if (act == B2_ACT_RELU) gz = (yv > 0.f) ? gz : 0.f;
        else if (act == B2_ACT_SIGMOID) gz = gz * ((1.f - yv) * yv);
      }
      if (kb == 0 && lane == 0) gb_acc += gz;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = kb + j * 32 + lane;
        if (k < K) {
          const float xv = __ldg(x + m * K + k);
          if (gx != nullptr) {
            float val = gz * __ldg(w + k);
            if (prev_act == B2_ACT_RELU) val = (xv > 0.f) ? val : 0.f;
            else if (prev_act == B2_ACT_SIGMOID) val = val * ((1.f - xv) * xv);
            gx[m * K + k] = val;
            if (gx_small != nullptr) gx_small[m * K + k] = tf32_small(val);
            cs[j] += val;
          }
          part[j] = fmaf(gz, xv, part[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kb + j * 32 + lane;
      if (k < K) {
        atomicAdd(sgw + k, part[j]);
        if (gb_prev != nullptr) atomicAdd(sgp + k, cs[j]);
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    if (sgw[k] != 0.f) b2_red_add(gw + k, sgw[k]);
    if (gb_prev != nullptr && sgp[k] != 0.f) b2_red_add(gb_prev + k, sgp[k]);
  }
  if (gb != nullptr) {
    const float t = b2_block_sum(gb_acc, red);
    if (threadIdx.x == 0 && t != 0.f) b2_red_add(gb, t);
  }
}
}  // namespace tc

// ---------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------
typedef CUresult (*b2_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                       const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static b2_encode_tiled_fn b2_get_encode() {
  static b2_encode_tiled_fn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<b2_encode_tiled_fn>(ptr);
  }
  return fn;
}

// K-major operand: memory (rows, K), K contiguous, leading dimension ld; box = 128 B of k x box_rows rows.
// MN-major operand: memory (K, rows), rows contiguous, leading dimension ld; box = 128 B of rows x (128/esz) k.
static int encode_operand(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld,
                          int mn_major, int box_rows, int esz) {
  b2_encode_tiled_fn enc = b2_get_encode();
  if (enc == nullptr) return b2_fail(B2_E_CUDA, "cuTensorMapEncodeTiled is unavailable in this driver");
  const cuuint32_t per128 = (cuuint32_t) (128 / esz);
  cuuint64_t dims[2], strides[1] = {(cuuint64_t) ld * (cuuint64_t) esz};
  cuuint32_t box[2], estr[2] = {1, 1};
  if (!mn_major) {
    dims[0] = (cuuint64_t) K; dims[1] = (cuuint64_t) rows;
    box[0] = per128; box[1] = (cuuint32_t) box_rows;
  } else {
    dims[0] = (cuuint64_t) rows; dims[1] = (cuuint64_t) K;
    box[0] = per128; box[1] = per128;
  }
  CUresult r = enc(map, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (mn_major && esz == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return b2_fail(B2_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int) r);
  return B2_OK;
}

static bool tma_ok_e(const void* p, int64_t ld, int esz) {
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && ((ld * esz) % 16 == 0);
}
static bool tma_ok(const float* p, int64_t ld) { return tma_ok_e(p, ld, 4); }

extern "C" B2_API int b2_gemm_tc_supported(const float* a, int64_t lda, const float* b, int64_t ldb,
                                           int64_t M, int64_t N, int64_t K) {
  return (tma_ok(a, lda) && tma_ok(b, ldb) && M >= 1 && N >= 1 && K >= 1 && M < (1ll << 31) &&
          N < (1ll << 31) && K < (1ll << 31)) ? 1 : 0;
}

// plan != NULL: fill in the launch plan (tile shape, ring depth, TMEM, shared memory) and return without
// touching the device — pure host arithmetic, so the CPU test-suite can sweep it (tests/test_abi.py).
static int gemm_tc_impl(const b2_gemm_desc* d, void* stream, b2_gemm_plan* plan) {
  B2_REQUIRE(d != nullptr, "NULL descriptor");
  const void* a = d->a; const void* b = d->b; float* c = d->c;
  const int64_t M = d->M, N = d->N, K = d->K, lda = d->lda, ldb = d->ldb, ldc = d->ldc;
  B2_REQUIRE(a && b && c, "NULL operand");
  B2_REQUIRE(d->elem_dtype == B2_F32 || d->elem_dtype == B2_BF16, "operand dtype must be B2_F32 or B2_BF16");
  const int esz = (d->elem_dtype == B2_BF16) ? 2 : 4;
  B2_REQUIRE(M >= 1 && N >= 1 && K >= 1 && ldc >= N, "bad shape");
  B2_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "shape exceeds 31 bits");
  B2_REQUIRE(lda >= (d->a_mn_major ? M : K) && ldb >= (d->b_mn_major ? N : K), "leading dimension too small");
  B2_REQUIRE(d->act >= B2_ACT_NONE && d->act <= B2_ACT_SIGMOID, "bad activation code %d", d->act);
  B2_REQUIRE(d->act_bwd >= B2_ACT_NONE && d->act_bwd <= B2_ACT_SIGMOID, "bad act_bwd code %d", d->act_bwd);
  B2_REQUIRE(d->act_bwd == B2_ACT_NONE || d->ybwd != nullptr, "act_bwd needs ybwd");
  B2_REQUIRE((d->a_small == nullptr) == (d->b_small == nullptr), "3xTF32 needs both small operands");
  B2_REQUIRE(esz == 4 || d->a_small == nullptr, "bf16 operands are single-pass (no small parts)");
  const bool inline_split = (d->flags & B2_GEMM_X3_INLINE) != 0;
  B2_REQUIRE(!inline_split || (esz == 4 && d->a_small == nullptr), "B2_GEMM_X3_INLINE: fp32 operands, no small parts");
  const bool three_pass = inline_split || d->a_small != nullptr;
  const int64_t ld_aux = d->ld_aux > 0 ? d->ld_aux : ldc;
  B2_REQUIRE(d->c_small == nullptr || ld_aux >= N, "ld_aux too small");
  if (!tma_ok_e(a, lda, esz) || !tma_ok_e(b, ldb, esz) ||
      (d->a_small != nullptr && !(tma_ok(d->a_small, lda) && tma_ok(d->b_small, ldb))))
    return b2_fail(B2_E_UNSUPPORTED, "operands are not TMA-addressable (16-byte base, 16-byte row pitch)");
  cudaStream_t st = (cudaStream_t) stream;

  // Tile-shape choice: BN in {32..256 step 32}; the kernel is bound by L2->SM operand traffic,
  // so minimise waves x per-CTA operand bytes, where one wave = 148 CTAs (one per SM).
  const int bke = 128 / esz;
  const int64_t tiles_m = b2_ceil_div(M, tc::BM);
  const int64_t num_kb = b2_ceil_div(K, bke);
  // split-K adds partial tiles with red.global: only for a plain linear epilogue
  const bool linear = (d->act == B2_ACT_NONE && d->mul == nullptr && d->add == nullptr && d->ybwd == nullptr &&
                       d->c_small == nullptr && d->c_pre == nullptr && d->colsum == nullptr);
  int best_bn = 0, best_split = 1;
  double best_cost = 1e300;
  // 3xTF32: (chains + 1 correction range) x bn <= 512 TMEM columns and 3 stages of {A, As, B, Bs} <= 227 KB:
  // bn <= 128 keeps 3-4 chains, bn = 160 keeps 2 (used when it saves a whole wave, e.g. 8192 x 624 x 624)
  const int x3_bn_max = [] { const char* e = getenv("B2_X3_BN_MAX"); return e ? atoi(e) : 160; }();
  const int bn_max = three_pass ? x3_bn_max : 256;
  // cycles per k-block (128 bytes of K): the tensor pipe needs passes x 4 instructions x bn/2, the L2->SM
  // fabric (~6300 B/cycle chip-wide, ~43 per SM) needs the operand bytes — 3xTF32 with small parts from HBM
  // moves them twice
  const double mma_per_bn = (three_pass ? 3.0 : 1.0) * 2.0;
  const double l2_per_row = 3.0 * ((three_pass && !inline_split) ? 2.0 : 1.0);
  const int bn_step = (d->b_mn_major && esz == 2) ? 64 : 32;   // an MN-major box is 128 bytes of rows
  for (int bn = bn_step; bn <= bn_max; bn += bn_step) {
    const int64_t tiles_n = b2_ceil_div(N, bn);
    for (int split = 1; split <= 32; split *= 2) {
      if (split > 1 && (!linear || num_kb / split < 8)) break;
      const int64_t ctas = tiles_m * tiles_n * split;
      const int64_t waves = b2_ceil_div(ctas, B2_NUM_SMS);
      const double kb = (double) b2_ceil_div(num_kb, split);
      // per-CTA time ~ fixed prologue/epilogue + k-blocks x bytes per k-block
      const double per_kb = fmax(mma_per_bn * bn, l2_per_row * (double) (tc::BM + bn));
      const double cost = (double) waves * (10000.0 + kb * per_kb);
      if (cost < best_cost) { best_cost = cost; best_bn = bn; best_split = split; }
    }
  }
  tc::Params p;
  const int nseg = three_pass ? 3 : 1;
  const void* as[2] = {a, d->a_small};
  const void* bs[2] = {b, d->b_small};
  for (int s = 0; plan == nullptr && s < (d->a_small != nullptr ? 2 : 1); ++s) {
    int rc = encode_operand(&p.map_a[s], as[s], M, K, lda, d->a_mn_major, tc::BM, esz);
    if (rc != B2_OK) return rc;
    rc = encode_operand(&p.map_b[s], bs[s], N, K, ldb, d->b_mn_major, best_bn, esz);
    if (rc != B2_OK) return rc;
  }
  // TMA-store epilogue for a plain C output (no split-K reduction, no accumulate, no second output):
  // C is described as an (N, M) fp32 tensor with pitch ldc and written in 32 x 32 boxes
  p.tma_store = 0;
  if (best_split == 1 && !d->beta_accumulate && d->c_small == nullptr && d->c_pre == nullptr && tma_ok(c, ldc)) {
    b2_encode_tiled_fn enc = plan != nullptr ? nullptr : b2_get_encode();
    cuuint64_t dims[2] = {(cuuint64_t) N, (cuuint64_t) M}, strides[1] = {(cuuint64_t) ldc * 4};
    cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
    if (plan != nullptr) p.tma_store = 1;
    if (enc != nullptr &&
        enc(&p.map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
      p.tma_store = 1;
  }
  {
    static const bool off = [] { const char* e = getenv("B2_GEMM_TMA_STORE"); return e != nullptr && atoi(e) == 0; }();
    if (off) p.tma_store = 0;
  }
  p.c = c; p.c_small = reinterpret_cast<float*>(d->c_small); p.c_pre = d->c_pre; p.ldc = ldc; p.ld_aux = ld_aux;
  p.bias = d->bias; p.mul = d->mul; p.add = d->add;
  p.ybwd = d->ybwd; p.colsum = d->colsum;
  p.M = (int) M; p.N = (int) N; p.K = (int) K; p.bn = best_bn; p.nseg = nseg; p.inline_split = inline_split ? 1 : 0; p.act = d->act;
  p.act_bwd = d->act_bwd; p.esz = esz;
  p.a_mn = d->a_mn_major ? 1 : 0; p.b_mn = d->b_mn_major ? 1 : 0;
  p.beta = d->beta_accumulate ? 1 : 0;
  p.kb_per_split = (int) b2_ceil_div(num_kb, best_split);
  const int splits = (int) b2_ceil_div(num_kb, p.kb_per_split);
  const int64_t tiles_n = b2_ceil_div(N, best_bn);
  const int64_t total_tiles = tiles_m * tiles_n * splits;
  B2_REQUIRE(total_tiles < (1ll << 31), "too many tiles");
  const bool multi = total_tiles > B2_NUM_SMS;        // some CTA processes more than one tile
  if (nseg == 1) {
    p.nmain = 1;
  } else {
    // Accumulation chains of the main product: the truncating accumulator loses accuracy with the LENGTH of
    // a chain, so the K range is cut into as many chains as TMEM holds (<= 4).  Fewer, longer chains were
    // measured (B2_X3_CHAIN_KB = k-blocks per chain): 0.003 ms faster per step, but 64 per chain gives
    // 2.1-2.8e-6 vs fp64 at K = 8,192 and even 16 per chain breaks the 1e-5 MLP-chain parity at 500-wide
    // layers — the default (0) keeps every chain TMEM has room for.
    static const int chain_kb = [] { const char* e = getenv("B2_X3_CHAIN_KB"); return e ? atoi(e) : 0; }();
    const int fit = 512 / best_bn - 1;                 // all of TMEM: nmain main chains + 1 correction
    p.nmain = chain_kb > 0 ? (int) b2_ceil_div(p.kb_per_split, chain_kb) : fit;
    if (p.nmain > fit) p.nmain = fit;
    if (p.nmain > 4) p.nmain = 4;
    if (p.nmain > p.kb_per_split) p.nmain = p.kb_per_split;
    if (p.nmain < 1) p.nmain = 1;
  }
  // Two accumulator stages (the epilogue of tile j under the main loop of tile j+1) when a CTA has
  // several tiles and both stages fit the 512 TMEM columns; 3xTF32 gives up chains for it only down to 2
  // (or to the k-block count of a short contraction).
  p.nacc = 1;
  if (multi) {
    int nm = p.nmain;
    const int corr = nseg > 1 ? 1 : 0;
    while (nm > 1 && 2 * (nm + corr) * best_bn > 512) --nm;
    const int floor_nm = (nseg > 1) ? (p.nmain < 2 ? p.nmain : 2) : 1;
    if (2 * (nm + corr) * best_bn <= 512 && nm >= floor_nm) { p.nmain = nm; p.nacc = 2; }
  }
  {
    const int need = p.nacc * (p.nmain + (nseg > 1 ? 1 : 0)) * best_bn;
    int cols = 32;
    while (cols < need) cols <<= 1;
    p.tmem_cols = cols;
  }
  p.tiles_m = (int) tiles_m; p.tiles_n = (int) tiles_n; p.splits = splits;
  if (plan == nullptr && splits > 1 && !p.beta && !(d->flags & B2_GEMM_C_IS_ZERO)) {
    cudaError_t e = cudaMemset2DAsync(c, (size_t) ldc * 4, 0, (size_t) N * 4, (size_t) M, st);
    if (e != cudaSuccess) return b2_fail(B2_E_CUDA, "b2_gemm_tc: memset: %s", cudaGetErrorString(e));
  }
  if (plan == nullptr && d->colsum != nullptr && !(d->flags & B2_GEMM_COLSUM_IS_ZERO)) {
    cudaError_t e = cudaMemsetAsync(d->colsum, 0, sizeof(float) * (size_t) N, st);
    if (e != cudaSuccess) return b2_fail(B2_E_CUDA, "b2_gemm_tc: memset: %s", cudaGetErrorString(e));
  }
  // operand ring: 4 stages of {A, B}, or 3 of {A, As, B, Bs} for 3xTF32 — one fewer when the widest
  // 3xTF32 tile (bn = 160) would not leave room for the epilogue patches under 227 KB
  const size_t stage_bytes = (nseg > 1 ? (size_t) 2 : (size_t) 1) * (tc::A_BYTES + (size_t) best_bn * 128);
  const int grid = (int) (total_tiles < B2_NUM_SMS ? total_tiles : B2_NUM_SMS);     // persistent: at most one CTA per SM
  // (one tile per CTA: the epilogue patches reuse the ring, see the kernel)
  const size_t fixed_bytes = (total_tiles <= grid ? 0 : tc::PATCH_BYTES) + 1024 + 192;
  p.dbg = 0; p.trace = nullptr;
#ifdef B2_GEMM_PROBE
  { const char* e = getenv("B2_GEMM_DBG"); p.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("B2_GEMM_TRACE"); p.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
#endif
  p.stages = tc::MAX_STAGES;       // as deep as 227 KB allows (3xTF32 stages are twice the size)
  if (p.dbg & 8) p.stages = 2;
  while (p.stages > 2 && p.stages * stage_bytes + fixed_bytes > (size_t) 227 * 1024) --p.stages;
  B2_REQUIRE(p.stages * stage_bytes + fixed_bytes <= (size_t) 227 * 1024, "tile does not fit shared memory");
  B2_REQUIRE(p.stages * stage_bytes >= (size_t) tc::PATCH_BYTES, "ring smaller than the epilogue patches");
  const size_t smem = p.stages * stage_bytes + fixed_bytes;
  if (plan != nullptr) {
    plan->bn = p.bn; plan->splits = p.splits; plan->stages = p.stages; plan->nacc = p.nacc; plan->nmain = p.nmain;
    plan->tmem_cols = p.tmem_cols; plan->grid = grid; plan->threads = tc::NTHREADS; plan->tiles_m = p.tiles_m;
    plan->tiles_n = p.tiles_n; plan->tma_store = p.tma_store; plan->passes = nseg; plan->kb_per_split = p.kb_per_split;
    plan->smem_bytes = (int64_t) smem;
    return B2_OK;
  }
  // opt-in to > 48 KB of dynamic shared memory: an idempotent per-process property of the kernel
  // (C++11 guarantees the initialiser runs once, thread-safely)
  static const cudaError_t attr_rc = cudaFuncSetAttribute(
      tc::gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (attr_rc != cudaSuccess) return b2_fail(B2_E_CUDA, "b2_gemm_tc: smem attribute: %s", cudaGetErrorString(attr_rc));
  B2_LAUNCH(tc::gemm_tf32_kernel, grid, tc::NTHREADS, smem, st, p);
  B2_CUDA_LAUNCH_CHECK("b2_gemm_tc");
  return B2_OK;
}

extern "C" B2_API int b2_gemm_tc_ex(const b2_gemm_desc* d, void* stream) { return gemm_tc_impl(d, stream, nullptr); }

extern "C" B2_API int b2_gemm_tc_plan(const b2_gemm_desc* d, b2_gemm_plan* plan) {
  B2_REQUIRE(plan != nullptr, "NULL plan");
  return gemm_tc_impl(d, nullptr, plan);
}

// fp32 (rows, cols; ld_in) -> bf16 (rows, cols; ld_out), round-to-nearest-even: the bf16-mode operand of a
// tensor whose producer is not one of our epilogues (weights once per step, the first layer's input).
namespace tc {
__global__ void __launch_bounds__(256)
to_bf16_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld_in,
               __nv_bfloat16* __restrict__ out, int64_t ld_out) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    out[r * ld_out + c] = __float2bfloat16_rn(__ldg(x + r * ld_in + c));
  }
}
}  // namespace tc

extern "C" B2_API int b2_to_bf16(const float* x, int64_t rows, int64_t cols, int64_t ld_in, void* out,
                                 int64_t ld_out, void* stream) {
  B2_REQUIRE(x && out && rows >= 0 && cols >= 0 && ld_in >= cols && ld_out >= cols, "bad argument");
  if (rows == 0 || cols == 0) return B2_OK;
  int64_t blocks = b2_ceil_div(rows * cols, 256);
  if (blocks > (int64_t) B2_NUM_SMS * 8) blocks = (int64_t) B2_NUM_SMS * 8;
  tc::to_bf16_kernel<<<(int) blocks, 256, 0, (cudaStream_t) stream>>>(x, rows, cols, ld_in,
                                                                       reinterpret_cast<__nv_bfloat16*>(out), ld_out);
  B2_CUDA_LAUNCH_CHECK("b2_to_bf16");
  return B2_OK;
}

extern "C" B2_API int b2_gemm_tc(const float* a, int64_t lda, const float* b, int64_t ldb, float* c,
                                 int64_t ldc, int64_t M, int64_t N, int64_t K, const float* bias, int act,
                                 const float* mul, const float* add, int beta_accumulate,
                                 const float* a_small, const float* b_small, void* stream) {
  b2_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = a; d.b = b; d.a_small = a_small; d.b_small = b_small; d.c = c;
  d.bias = bias; d.mul = mul; d.add = add;
  d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.M = M; d.N = N; d.K = K;
  d.act = act; d.beta_accumulate = beta_accumulate;
  return b2_gemm_tc_ex(&d, stream);
}

extern "C" B2_API int b2_split_tf32(const float* x, float* small, int64_t n, void* stream) {
  B2_REQUIRE(x && small, "NULL pointer");
  if (n <= 0) return B2_OK;
  int64_t blocks = b2_ceil_div(n, 256);
  if (blocks > (int64_t) B2_NUM_SMS * 8) blocks = (int64_t) B2_NUM_SMS * 8;
  B2_LAUNCH(tc::split_tf32_kernel, (int) blocks, 256, 0, (cudaStream_t) stream, x, small, n);
  B2_CUDA_LAUNCH_CHECK("b2_split_tf32");
  return B2_OK;
}

extern "C" B2_API int b2_transpose_f32(const float* in, int64_t rows, int64_t cols, int64_t ld_in,
                                       float* out, int64_t ld_out, float* out_small, void* stream) {
  B2_REQUIRE(in && out, "NULL pointer");
  B2_REQUIRE(rows >= 0 && cols >= 0 && ld_in >= cols && ld_out >= rows, "bad shape");
  if (rows == 0 || cols == 0) return B2_OK;
  dim3 grid((unsigned) b2_ceil_div(cols, 32), (unsigned) b2_ceil_div(rows, 32));
  B2_REQUIRE(grid.y <= 65535, "too many rows for this launch geometry");
  tc::transpose_kernel<<<grid, 256, 0, (cudaStream_t) stream>>>(in, rows, cols, ld_in, out, ld_out, out_small);
  B2_CUDA_LAUNCH_CHECK("b2_transpose_f32");
  return B2_OK;
}

extern "C" B2_API int b2_prep_operand(const float* x, const float* y, int act, int64_t R, int64_t C,
                                      float* out, float* out_small, float* outT, float* outT_small,
                                      float* colsum, void* stream) {
  B2_REQUIRE(x != nullptr, "NULL input");
  B2_REQUIRE(R >= 0 && C >= 0, "bad shape");
  B2_REQUIRE((act >= B2_ACT_NONE && act <= B2_ACT_SIGMOID) || act == B2_PREP_MUL, "bad activation code %d", act);
  cudaStream_t st = (cudaStream_t) stream;
  if (colsum != nullptr) {
    cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(float) * (size_t) C, st);
    if (e != cudaSuccess) return b2_fail(B2_E_CUDA, "b2_prep_operand: memset: %s", cudaGetErrorString(e));
  }
  if (R == 0 || C == 0) return B2_OK;
  dim3 grid((unsigned) b2_ceil_div(C, 32), (unsigned) b2_ceil_div(R, 32));
  B2_REQUIRE(grid.y <= 65535, "too many rows for this launch geometry");
  B2_LAUNCH(tc::prep_operand_kernel, grid, 256, 0, st, x, y, act, R, C, C, out, out_small, outT, outT_small, colsum);
  B2_CUDA_LAUNCH_CHECK("b2_prep_operand");
  return B2_OK;
}

extern "C" B2_API int b2_head_fwd(const float* x, const float* w, const float* b, int64_t M, int K, int act,
                                  float* y, void* stream) {
  B2_REQUIRE(x && w && y, "NULL pointer");
  B2_REQUIRE(K >= 1 && act >= B2_ACT_NONE && act <= B2_ACT_SIGMOID, "bad K/act");
  if (M <= 0) return B2_OK;
  int64_t blocks = b2_ceil_div(M * 32, 256);
  if (blocks > (int64_t) B2_NUM_SMS * 8) blocks = (int64_t) B2_NUM_SMS * 8;
  B2_LAUNCH(tc::head_fwd_kernel, (int) blocks, 256, 0, (cudaStream_t) stream, x, w, b, M, K, act, y);
  B2_CUDA_LAUNCH_CHECK("b2_head_fwd");
  return B2_OK;
}

extern "C" B2_API int b2_head_bwd(const float* x, const float* w, const float* y, const float* gy, int64_t M,
                                  int K, int act, float* gx, float* gw, float* gb, void* stream) {
  return b2_head_bwd_ex(x, w, y, gy, M, K, act, gx, gw, gb, B2_ACT_NONE, nullptr, nullptr, 0, stream);
}

extern "C" B2_API int b2_head_bwd_ex(const float* x, const float* w, const float* y, const float* gy, int64_t M,
                                     int K, int act, float* gx, float* gw, float* gb, int prev_act,
                                     float* gx_small, float* gb_prev, int grads_zeroed, void* stream) {
  B2_REQUIRE(x && w && gy && gw, "NULL pointer");
  B2_REQUIRE(K >= 1 && K <= 6144 && act >= B2_ACT_NONE && act <= B2_ACT_SIGMOID, "bad K/act");
  B2_REQUIRE(prev_act >= B2_ACT_NONE && prev_act <= B2_ACT_SIGMOID, "bad prev_act");
  B2_REQUIRE(act == B2_ACT_NONE || y != nullptr, "activation backward needs y");
  B2_REQUIRE(gx != nullptr || (gx_small == nullptr && gb_prev == nullptr && prev_act == B2_ACT_NONE),
             "prev_act / gx_small / gb_prev need gx");
  cudaStream_t st = (cudaStream_t) stream;
  cudaError_t e = cudaSuccess;
  if (!grads_zeroed) {    // the caller vouches that gw / gb / gb_prev are all-zero (a gradient arena cleared by Adam)
    e = cudaMemsetAsync(gw, 0, sizeof(float) * (size_t) K, st);
    if (e == cudaSuccess && gb_prev != nullptr) e = cudaMemsetAsync(gb_prev, 0, sizeof(float) * (size_t) K, st);
    if (e == cudaSuccess && gb != nullptr) e = cudaMemsetAsync(gb, 0, sizeof(float), st);
  }
  if (e != cudaSuccess) return b2_fail(B2_E_CUDA, "b2_head_bwd: memset: %s", cudaGetErrorString(e));
  if (M <= 0) return B2_OK;
  int64_t ctas = 2 * B2_NUM_SMS;
  if (ctas > b2_ceil_div(M, 8)) ctas = b2_ceil_div(M, 8);
  const int64_t rows_per_cta = b2_ceil_div(M, ctas);
  ctas = b2_ceil_div(M, rows_per_cta);
  const float* y_arg = (act == B2_ACT_NONE) ? nullptr : y;
  B2_LAUNCH(tc::head_bwd_kernel, (int) ctas, 256, 2 * sizeof(float) * (size_t) K, st,
            x, w, y_arg, gy, M, K, act, rows_per_cta, gx, gw, gb, prev_act, gx_small, gb_prev);
  B2_CUDA_LAUNCH_CHECK("b2_head_bwd");
  return B2_OK;
}
