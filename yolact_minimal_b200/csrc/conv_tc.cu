// tcgen05 implicit-GEMM convolution for sm_100a (fp16 / bf16 operands, fp32 accumulation in TMEM).
//
//   out[M, Cout] = epilogue( sum_{tap} A[M + shift_tap, Cin] * W_tap[Cin, Cout] )
//
// on the haloed NHWC layout of layers.cuh, where every convolution tap is a constant row shift
// of the activation matrix.  Persistent, warp-specialised CTA (320 threads, 1 CTA / SM):
//   warp 0      TMA producer: per k-block one 128x64 A tile (row coordinate m0 + shift_tap,
//               out-of-range rows zero-filled by TMA = the conv padding) and one BNx64 weight
//               tile, both 128B-swizzled, into a ring of shared-memory stages (mbarrier full/empty);
//               small-K layers keep their whole weight slice resident instead
//   warp 1      allocates TMEM, issues tcgen05.mma (M=128, N=BN, K=16) from one elected lane,
//               tcgen05.commit releases smem stages and publishes finished accumulators
//   warps 2..9  eight independent epilogue warps (two per TMEM lane quadrant, alternating 32-column
//               chunks): tcgen05.ld of the fp32 accumulator (double-buffered in TMEM so the next tile's
//               MMAs overlap), + folded-BN bias, + residual, ReLU/clamp/pack, halo zeroing, swizzled
//               smem staging and one TMA store per warp and chunk (or dense fp32 stores)
// Tiles are scheduled round-robin over the persistent grid, N-tiles of one M-tile adjacent so
// the A tile is shared through L2.
//
// Per-layer forms chosen in tc_plan_create (measurements: profiles/r1_conv_tc_experiments.txt):
//   PAIR      cluster of two CTAs, tcgen05 cta_group::2: one 256 x BN tile per pair, each CTA stages its
//             own 128 A rows and half of the weight tile; the leader (rank 0) issues the MMAs
//   slab      3x3 stride 1: one [130 x 64] A slab per (tap row, k-block) serves the three dx taps through
//             row-shifted shared-memory descriptors; slab ring and weight ring advance separately
//   res_kb    residual added by the tensor core: BN/64 extra k-blocks against an identity tile
//   alt_tiles BN <= 64: the two epilogue groups take alternate tiles (one accumulator buffer each) instead of alternate chunks
// Launches use programmatic dependent launch: everything before griddep_wait() touches only this
// CTA's shared memory / TMEM (and the constant weights), so it overlaps the previous kernel's tail.
//
// Environment switches (tooling / A-B runs only): YOLACT_B200_PAIR=0|1, YOLACT_B200_RESMMA=0|1,
// YOLACT_B200_NO_SLAB, YOLACT_B200_NO_ALT, YOLACT_B200_NO_BRES, YOLACT_B200_NO_PDL, YOLACT_B200_BN=<n>, YOLACT_B200_NRES=<n>.
#include "layers.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

namespace yb {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                 // bf16 elements = 128 bytes = one swizzle row
constexpr int TC_THREADS = 320;             // TMA warp, MMA warp, 2 x 4 epilogue warps
constexpr int TC_OUT_BUFS = 2;                // output staging buffers per epilogue warp
constexpr uint32_t TC_WARP_TILE = 32 * 32 * 2;  // 2 KiB: one warp's 32 rows x 32 columns of 16-bit outputs
constexpr uint32_t TC_A_STAGE = TC_BM * TC_BK * 2;   // 16 KiB
constexpr uint32_t TC_SLAB_ROWS = TC_BM + 2;         // slab mode: rows m0-1 .. m0+128 of one tap row
constexpr uint32_t TC_A_SLAB = 136 * TC_BK * 2;      // 17 KiB slot: the 130-row slab padded to the 1 KiB swizzle period
constexpr uint32_t TC_S16_ROWS = TC_BM + 3;          // stem16: pixels p-1 .. p+129 of one tap row serve dx = 0 .. 3
constexpr uint32_t TC_S16_SLOT = 5120;               // 131 rows x 32 B, padded to 1 KiB

struct TcParams {
  long long M;            // rows to produce (B * plane)
  int m_tiles, n_tiles;
  int kb_per_tap;         // Cin / 64
  int ntaps;
  int tap_shift[kMaxTaps];
  int BN;                 // N tile (multiple of 16, <= 256)
  int tmem_cols;          // allocated columns (power of two >= 2*BN)
  int stages;
  int Cout, Cout_pad, relu, out_mode;
  int is_f16;             // operands fp16 (else bf16)
  int tma_epi;            // out_mode 0 && BN % 32 == 0: smem-staged TMA stores (+ TMA-prefetched residual)
  int nres;               // residual prefetch buffers (0 if none)
  int b_resident;         // weights of the CTA's N tile stay in shared memory for all its M tiles
  int slab, stages_a;     // slab: 3x3 stride-1 convs load ONE [130 x 64] A slab per (tap row, k-block) and run the three dx taps off
                          // it through descriptors that start 0 / 128 / 256 B into the slab (UMMA swizzles on absolute shared-memory
                          // address bits, so a row-shifted start needs no base offset: probed once, and every 3x3 parity test runs through it).  A third of the A
                          // traffic; the slab ring (stages_a slots of 17 KB) and the weight ring (stages) then advance separately.
  int alt_tiles;          // narrow tiles (BN <= 64): the two epilogue groups take alternate TILES (group g owns accumulator g) instead of
                          // alternate 32-column chunks of every tile, halving the per-tile hand-offs each warp sits through
  int res_kb;             // > 0: the residual is added BY THE TENSOR CORE: BN/64 extra k-blocks whose A tile is the
                          // residual's [128 x 64] slab and whose B tile is a shared-memory 64x64 identity (N=64 MMAs
                          // into the matching 64 accumulator columns) -- the epilogue never touches the residual
  int stem16;             // stem form: A rows are single s2d pixels (16 channels = 32 bytes, SWIZZLE_32B); one [131 x 32 B] slab per tap row dy
                          // serves the four dx taps through descriptors that start 0 / 32 / 64 / 96 B into it (K = 16 per MMA), the 16 weight
                          // tiles [64 x 32 B] stay resident.  A quarter of the L2 -> shared-memory fill of the overlapping-rows form.
  int gemm;               // plain GEMM over long K (weight gradients): out[M x taps*Nper] (+)= A[M x K] * B_tap[Nper x K]^T with both operands
                          // K-major; N tile n -> tap n / gemm_ntile_tap, B rows (n % gemm_ntile_tap) * BN, B columns k + gemm_shift[tap]
                          // (the conv tap as a COLUMN offset of the transposed activations; out-of-range columns read as zero)
  int gemm_ntile_tap;     // N tiles per tap
  int gemm_shift[16];     // per-tap column shift of the B operand
  int accumulate;         // out_mode 2: out += result (shared weights: one launch per use)
  int gemm_splits;        // split-K: the K range of every output tile is cut into gemm_splits pieces of gemm_kb_split k-blocks, one
  int gemm_kb_split;      // scheduling unit each; the partial tiles are added to `out` with atomics (out is zeroed by the caller)
  Geom g;
  const float* bias;
  const void* residual;
  void* out;
};

struct TcPlan {
  CUtensorMap tmA, tmB, tmOut, tmRes;
  int BN, stages, tmem_cols, tma_epi, nres, b_resident, res_kb, pair, slab, stages_a, alt_tiles, stem16;
  int sms;                 // SM count of the device the plan was created on
  int gemm;                // plain-GEMM plan (tc_plan_create_gemm)
  size_t smem_bytes;
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
// bring one box of a tensor into L2 (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// programmatic dependent launch: let the next grid's CTAs take over SMs as ours retire / block until the previous grid's
// memory is complete and visible (both are no-ops for a launch without the programmatic-serialization attribute)
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// ---- CTA pair (cluster of 2, tcgen05 cta_group::2) helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default (.release.cta) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id): the TMEM hand-off is ordered by the
  // tcgen05 fences on both sides; a .release.cluster arrive costs a MEMBAR.ALL.GPU per epilogue warp and tile
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
// completion of all prior MMAs of the pair -> the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// K-major, 32B-swizzled operand tile (stem16): rows of 32 bytes (K = 16), 8-row groups 256 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;                             // SWIZZLE_32B
  return d;
}
// MN-major, 128B-swizzled operand tile (the weight-gradient GEMM's B operand, read straight from the [pixels][channels] activations):
// the tile is BN/64 TMA boxes of [64 k-rows][64 channels] (8 KiB each); a row of 128 bytes holds 64 consecutive MN elements of one k,
// 8 k-rows form a 1 KiB swizzle atom (stride byte offset 1024), the next 64-element MN chunk is the next box (leading byte offset 8192).
// Canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address
  d |= (uint64_t)(8192 >> 4) << 16;                   // leading byte offset: between 64-element MN chunks
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: between groups of 8 k-rows
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// The MMA warp runs its loops with all 32 lanes (warp-uniform control flow and operands, so descriptors and addresses stay
// in uniform registers); only the tcgen05 instructions themselves are predicated on the one elected lane (`issue`).  Issued from
// inside an `if (lane == 0)` region instead, every MMA costs an ELECT / R2UR.BROADCAST / BRA.U.ANY sequence of ~15 dependent
// instructions -- ~130 cycles, which bounds the N = 64 / 128 tiles (32 / 64 tensor cycles per MMA).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <bool F16> __device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (F16) {
    __half2 v = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
// one F2FP per pair: round-to-nearest pack with the fp16 range clamp (and optionally ReLU) folded into the conversion
template <bool F16, bool RELU> __device__ __forceinline__ uint32_t pack2_sat(float a, float b) {
  uint32_t r;
  if (F16) {
    if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  } else {
    if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  }
  return r;
}
template <bool F16> __device__ __forceinline__ float2 unpack2(uint32_t u) {
  if (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}

// epilogue for `n` (16 or 32) consecutive columns starting at global column col of row m
template <int NCOL, bool F16>
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, const uint32_t* acc, long long m, int col, bool valid, bool halo,
                                               int img, int y, int x) {
  if (!valid) return;
  if (p.out_mode == 2) {                                           // plain fp32 row-major [M][Cout_pad], no bias / activation
    float* o = (float*)p.out + m * p.Cout_pad + col;
#pragma unroll
    for (int i = 0; i < NCOL; i += 4) {
      const float4 r = make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]), __uint_as_float(acc[i + 3]));
      if (p.accumulate) atomicAdd(reinterpret_cast<float4*>(o + i), r);         // split-K partial tiles / shared weights: 128-bit reduction
      else *reinterpret_cast<float4*>(o + i) = r;
    }
    return;
  }
  float v[NCOL];
#pragma unroll
  for (int i = 0; i < NCOL; i += 4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
    v[i] = __uint_as_float(acc[i]) + b.x; v[i + 1] = __uint_as_float(acc[i + 1]) + b.y;
    v[i + 2] = __uint_as_float(acc[i + 2]) + b.z; v[i + 3] = __uint_as_float(acc[i + 3]) + b.w;
  }
  if (p.out_mode == 0) {
    uint16_t* o = (uint16_t*)p.out + m * p.Cout + col;
    if (halo) {
#pragma unroll
      for (int i = 0; i < NCOL; i += 8) *reinterpret_cast<uint4*>(o + i) = make_uint4(0, 0, 0, 0);
      return;
    }
    if (p.residual) {
      const uint16_t* r = (const uint16_t*)p.residual + m * p.Cout + col;
#pragma unroll
      for (int i = 0; i < NCOL; i += 8) {
        const uint4 rv = __ldg(reinterpret_cast<const uint4*>(r + i));
        const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack2<F16>(rw[j]);
          v[i + 2 * j] += f.x; v[i + 2 * j + 1] += f.y;
        }
      }
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < NCOL; ++i) v[i] = apply_act(v[i], p.relu);
    }
#pragma unroll
    for (int i = 0; i < NCOL; i += 8) {
      uint4 pk;
      pk.x = pack2<F16>(v[i], v[i + 1]); pk.y = pack2<F16>(v[i + 2], v[i + 3]);
      pk.z = pack2<F16>(v[i + 4], v[i + 5]); pk.w = pack2<F16>(v[i + 6], v[i + 7]);
      *reinterpret_cast<uint4*>(o + i) = pk;
    }
  } else if (!halo) {
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < NCOL; ++i) v[i] = apply_act(v[i], p.relu);
    }
    const long long r = ((long long)img * p.g.H + (y - 1)) * p.g.W + (x - 1);
    float* o = (float*)p.out + r * p.Cout_pad + col;
#pragma unroll
    for (int i = 0; i < NCOL; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}

// CTA-local iteration i -> tile coordinates.  Streaming mode: tiles round-robin over the grid with
// the N tiles of one M tile adjacent (A shared through L2).  Weight-resident mode: the CTA owns one
// N tile for its whole life (its weights stay in shared memory) and strides over M tiles.
template <bool PAIR>
__device__ __forceinline__ bool tile_at(const TcParams& p, int i, int& m_tile, int& n_tile) {
  // scheduling unit: one CTA, or one CTA pair (its M tile is 256 rows: 128 per CTA)
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // p.m_tiles counts unit tiles; the returned m_tile is THIS CTA's 128-row tile
  if (p.b_resident) {
    n_tile = unit % p.n_tiles;
    m_tile = unit / p.n_tiles + i * (units / p.n_tiles);
    if (m_tile >= p.m_tiles) return false;
  } else {
    int tile = unit + i * units;
    if (p.gemm) {                                                  // split-K: the split index is recovered by gemm_split_at()
      if (tile >= p.m_tiles * p.n_tiles * p.gemm_splits) return false;
      tile %= p.m_tiles * p.n_tiles;
    } else if (tile >= p.m_tiles * p.n_tiles) return false;
    m_tile = tile / p.n_tiles;
    n_tile = tile - m_tile * p.n_tiles;
  }
  if (PAIR) m_tile = 2 * m_tile + (int)cluster_ctarank();
  return true;
}

// split-K GEMM: k-block range [kb0, kb1) of CTA-local iteration i
__device__ __forceinline__ void gemm_kb_range(const TcParams& p, int i, int& kb0, int& kb1) {
  const int split = ((int)blockIdx.x + i * (int)gridDim.x) / (p.m_tiles * p.n_tiles);
  kb0 = split * p.gemm_kb_split;
  kb1 = min(kb0 + p.gemm_kb_split, p.kb_per_tap);
}

template <bool F16, bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B needs 1024B alignment
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;                              // CTA pair: 0 = leader (issues the MMAs)
  const int bn_cta = PAIR ? p.BN / 2 : p.BN;                                        // weight rows held by this CTA
  const uint32_t b_stage = p.stem16 ? 2048u : (uint32_t)bn_cta * TC_BK * 2;         // stem16: 64 rows x 32 B per (dy, dx) weight tile
  const int num_kb = p.stem16 ? 16 : p.ntaps * p.kb_per_tap;
  uint8_t* sA = smem;                                                               // [stages][16 KiB]
  uint8_t* sB = smem + (p.stem16 ? (size_t)p.stages_a * TC_S16_SLOT : p.slab ? (size_t)p.stages_a * TC_A_SLAB : (size_t)p.stages * TC_A_STAGE);   // [stages | num_kb][BN x 128 B]
  uint8_t* sOut = sB + (size_t)(p.b_resident ? num_kb : p.stages) * b_stage;        // [8 warps][2][32 rows x 64 B], SWIZZLE_64B
  uint8_t* sRes = sOut + (p.tma_epi ? 8 * TC_OUT_BUFS * TC_WARP_TILE : 0);          // [8 warps][nres][32 rows x 64 B]
  uint8_t* sEye = sRes;                                                             // res_kb: [64][128 B] identity, 128B-swizzled
  uint64_t* full = reinterpret_cast<uint64_t*>(sRes + (p.res_kb ? (size_t)8192 : (size_t)8 * p.nres * TC_WARP_TILE));   // pair: 32 identity rows per CTA
  uint64_t* empty = full + 8;
  uint64_t* tmem_full = empty + 8;             // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint64_t* res_full = tmem_empty + 2;         // [8 warps][4]
  uint64_t* b_full = res_full + 32;            // [1]
  uint64_t* full_a = b_full + 1;               // [8] slab ring (slab mode)
  uint64_t* empty_a = full_a + 8;              // [8]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(empty_a + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < 32; ++i) mbar_init(&res_full[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], (PAIR ? 16u : 8u) >> (p.alt_tiles ? 1 : 0)); }   // pair: both CTAs' epilogue warps free the leader's accumulators
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (p.res_kb && warp >= 2) {
    // identity operand tile: row n holds 1.0 at element n; element k of a row lives in 16-byte chunk (k/8) ^ (n & 7)
    // (pair: the N=64 identity is split like every B operand -- this CTA holds rows 32*rank .. 32*rank+31)
    const int t = (int)threadIdx.x - 64;                             // 0..255: four threads per row, two chunks each
    const int n = t >> 2;                                            // local row
    const int kone = PAIR ? n + 32 * (int)rank : n;                  // column of the 1.0 in that row
    const uint32_t one = F16 ? 0x3C00u : 0x3F80u;
    if (!PAIR || n < 32) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = (t & 3) * 2 + q;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (c == (kone >> 3)) {
          const int e = kone & 7;
          const uint32_t w = (e & 1) ? (one << 16) : one;
          if ((e >> 1) == 0) v.x = w; else if ((e >> 1) == 1) v.y = w; else if ((e >> 1) == 2) v.z = w; else v.w = w;
        }
        *reinterpret_cast<uint4*>(sEye + n * 128 + ((c ^ (n & 7)) << 4)) = v;
      }
    }
    fence_async_smem();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();               // pair: the peer's barriers / TMEM are ready too
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Everything above touched only shared memory / TMEM of this CTA.  From here on the previous kernel's output is read
  // and buffers it may still be reading are overwritten: every role passes griddep_wait() first (the producer after
  // requesting the weights, which no kernel writes).
  griddep_launch_dependents();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int m_tile, n_tile;
      // pair: every load lands in the issuing CTA's own shared memory but is counted on the LEADER's barrier, which
      // expects both CTAs' bytes (the leader's MMAs read both shared memories); each CTA recycles a stage on its own
      // `empty` barrier (the MMA commit is multicast to both)
      const uint32_t bfull_addr = PAIR ? map_to_rank(smem_u32(b_full), 0) : 0u;
      if (p.stem16 && tile_at<PAIR>(p, 0, m_tile, n_tile)) {          // 16 weight tiles [64 x 16], k = dy*64 + dx*16
        mbar_expect_tx(b_full, 16u * b_stage);
        for (int j = 0; j < 16; ++j) tma_load_2d(sB + (size_t)j * b_stage, &tmB, j * 16, 0, b_full);
      } else if (p.b_resident && tile_at<PAIR>(p, 0, m_tile, n_tile)) {
        // the CTA's whole weight slice [bn_cta x Ktot], loaded once
        if (!PAIR || rank == 0) mbar_expect_tx(b_full, (uint32_t)num_kb * b_stage * (PAIR ? 2u : 1u));
        for (int kb = 0; kb < num_kb; ++kb) {
          if (PAIR) tma_load_2d_pair(sB + (size_t)kb * b_stage, &tmB, kb * TC_BK, n_tile * p.BN + (int)rank * bn_cta, bfull_addr);
          else tma_load_2d(sB + (size_t)kb * b_stage, &tmB, kb * TC_BK, n_tile * p.BN, b_full);
        }
      }
      griddep_wait();
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = (p.b_resident ? TC_A_STAGE : TC_A_STAGE + b_stage) * (PAIR ? 2u : 1u);
      int sa = 0; uint32_t phase_a = 0;
      for (int i = 0; p.stem16 && tile_at<PAIR>(p, i, m_tile, n_tile); ++i) {
        for (int dy = 0; dy < 4; ++dy) {                               // one [131 px x 16 ch] slab per tap row
          mbar_wait(&empty_a[sa], phase_a ^ 1);
          mbar_expect_tx(&full_a[sa], TC_S16_ROWS * 32);
          tma_load_2d(sA + (size_t)sa * TC_S16_SLOT, &tmA, 0, m_tile * TC_BM + p.tap_shift[dy], &full_a[sa]);
          if (++sa == p.stages_a) { sa = 0; phase_a ^= 1; }
        }
      }
      for (int i = 0; p.slab && tile_at<PAIR>(p, i, m_tile, n_tile); ++i) {
        // slab mode: per (tap row dy, k-block) one A slab, then the three dx weight tiles that consume it
        const int m0 = m_tile * TC_BM, n0 = n_tile * p.BN;
        for (int dy = 0; dy < 3; ++dy) {
          const int row = m0 + p.tap_shift[3 * dy];
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            mbar_wait(&empty_a[sa], phase_a ^ 1);
            if (PAIR) {
              if (rank == 0) mbar_expect_tx(&full_a[sa], 2 * TC_SLAB_ROWS * 128);
              tma_load_2d_pair(sA + (size_t)sa * TC_A_SLAB, &tmA, kb * TC_BK, row, map_to_rank(smem_u32(&full_a[sa]), 0));
            } else {
              mbar_expect_tx(&full_a[sa], TC_SLAB_ROWS * 128);
              tma_load_2d(sA + (size_t)sa * TC_A_SLAB, &tmA, kb * TC_BK, row, &full_a[sa]);
            }
            if (++sa == p.stages_a) { sa = 0; phase_a ^= 1; }
            for (int dx = 0; dx < 3 && !p.b_resident; ++dx) {
              const int kcol = ((3 * dy + dx) * p.kb_per_tap + kb) * TC_BK;
              mbar_wait(&empty[stage], phase ^ 1);
              if (PAIR) {
                if (rank == 0) mbar_expect_tx(&full[stage], 2 * b_stage);
                tma_load_2d_pair(sB + (size_t)stage * b_stage, &tmB, kcol, n0 + (int)rank * bn_cta, map_to_rank(smem_u32(&full[stage]), 0));
              } else {
                mbar_expect_tx(&full[stage], b_stage);
                tma_load_2d(sB + (size_t)stage * b_stage, &tmB, kcol, n0, &full[stage]);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      for (int i = 0; !p.slab && !p.stem16 && tile_at<PAIR>(p, i, m_tile, n_tile); ++i) {
        const int m0 = m_tile * TC_BM;
        int n0 = n_tile * p.BN, bcol0 = 0;
        if (p.gemm) {                                              // B tile: rows of the tap's operand, columns shifted by the tap
          const int tap = n_tile / p.gemm_ntile_tap;
          n0 = (n_tile - tap * p.gemm_ntile_tap) * p.BN;
          bcol0 = p.gemm_shift[tap];
        }
        int kb_lo = 0, kb_hi = p.kb_per_tap;
        if (p.gemm) gemm_kb_range(p, i, kb_lo, kb_hi);
        for (int t = 0; t < p.ntaps; ++t) {
          const int row = m0 + p.tap_shift[t];
          for (int kb = kb_lo; kb < kb_hi; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (PAIR) {
              const uint32_t full_addr = map_to_rank(smem_u32(&full[stage]), 0);
              if (rank == 0) mbar_expect_tx(&full[stage], tx);
              tma_load_2d_pair(sA + (size_t)stage * TC_A_STAGE, &tmA, kb * TC_BK, row, full_addr);
              if (!p.b_resident) tma_load_2d_pair(sB + (size_t)stage * b_stage, &tmB, (t * p.kb_per_tap + kb) * TC_BK, n0 + (int)rank * bn_cta, full_addr);
            } else {
              mbar_expect_tx(&full[stage], tx);
              tma_load_2d(sA + (size_t)stage * TC_A_STAGE, &tmA, kb * TC_BK, row, &full[stage]);
              if (p.gemm) {                                          // MN-major B: BN/64 boxes [64 k-rows][64 channels] at k-row kb*64 + tap shift
                for (int j = 0; j < p.BN / 64; ++j)
                  tma_load_2d(sB + (size_t)stage * b_stage + (size_t)j * 8192, &tmB, n0 + j * 64, kb * TC_BK + bcol0, &full[stage]);
              } else if (!p.b_resident) tma_load_2d(sB + (size_t)stage * b_stage, &tmB, (t * p.kb_per_tap + kb) * TC_BK, n0, &full[stage]);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        for (int j = 0; j < p.res_kb; ++j) {                     // residual slabs ride the same pipeline (A slot only)
          mbar_wait(&empty[stage], phase ^ 1);
          if (PAIR) {
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * TC_A_STAGE);
            tma_load_2d_pair(sA + (size_t)stage * TC_A_STAGE, &tmRes, n0 + j * TC_BK, m0, map_to_rank(smem_u32(&full[stage]), 0));
          } else {
            mbar_expect_tx(&full[stage], TC_A_STAGE);
            tma_load_2d(sA + (size_t)stage * TC_A_STAGE, &tmRes, n0 + j * TC_BK, m0, &full[stage]);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (pair: the leader CTA issues M=256 MMAs over both CTAs' operands) =================
    if (!PAIR || rank == 0) {
      const uint32_t issue = elect_one();
      // instruction descriptor: D=f32, A=B=bf16 (1) or f16 (0), K-major both, N=BN, M=128
      const uint32_t fmt = F16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((PAIR ? 2 * TC_BM : TC_BM) >> 4) << 24);
      const uint32_t idesc_eye = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)((PAIR ? 2 * TC_BM : TC_BM) >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int sa = 0; uint32_t phase_a = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int m_tile, n_tile;
      if (p.b_resident && tile_at<PAIR>(p, 0, m_tile, n_tile)) mbar_wait(b_full, 0);
      for (int i = 0; tile_at<PAIR>(p, i, m_tile, n_tile); ++i) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
        if (p.stem16) {
          for (int dy = 0; dy < 4; ++dy) {
            mbar_wait(&full_a[sa], phase_a);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA + (size_t)sa * TC_S16_SLOT);
#pragma unroll
            for (int dx = 0; dx < 4; ++dx)                                        // tap dx = slab rows dx .. dx+127, K = 16
              umma_bf16(d_tmem, umma_desc_sw32(a_addr + (uint32_t)dx * 32u), umma_desc_sw32(smem_u32(sB + (size_t)(dy * 4 + dx) * b_stage)), idesc,
                        (dy | dx) != 0 ? 1u : 0u, issue);
            umma_commit(&empty_a[sa], issue);
            if (++sa == p.stages_a) { sa = 0; phase_a ^= 1; }
          }
        }
        if (p.slab) {
          int kbi = 0;
          for (int dy = 0; dy < 3; ++dy) {
            for (int kb = 0; kb < p.kb_per_tap; ++kb) {
              mbar_wait(&full_a[sa], phase_a);
              tc_fence_after();
              const uint32_t a_addr = smem_u32(sA + (size_t)sa * TC_A_SLAB);
              for (int dx = 0; dx < 3; ++dx, ++kbi) {
                if (!p.b_resident) { mbar_wait(&full[stage], phase); tc_fence_after(); }
                const uint64_t da = umma_desc(a_addr + (uint32_t)dx * 128u);         // tap dx = slab rows dx .. dx+127
                const uint64_t db = umma_desc(smem_u32(sB + (size_t)(p.b_resident ? (3 * dy + dx) * p.kb_per_tap + kb : stage) * b_stage));
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                  if (PAIR) umma_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kbi | k) != 0 ? 1u : 0u, issue);
                  else umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kbi | k) != 0 ? 1u : 0u, issue);
                }
                if (!p.b_resident) {
                  if (PAIR) umma_commit_pair(&empty[stage], issue); else umma_commit(&empty[stage], issue);
                  if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
              }
              if (PAIR) umma_commit_pair(&empty_a[sa], issue); else umma_commit(&empty_a[sa], issue);   // slab consumed by all three taps
              if (++sa == p.stages_a) { sa = 0; phase_a ^= 1; }
            }
          }
        }
        int kb_lo = 0, kb_hi = num_kb;
        if (p.gemm) gemm_kb_range(p, i, kb_lo, kb_hi);
        for (int kb = kb_lo; kb < kb_hi && !p.slab && !p.stem16; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc(smem_u32(sA + (size_t)stage * TC_A_STAGE));
          const uint32_t b_addr = smem_u32(sB + (size_t)(p.b_resident ? kb : stage) * b_stage);
          const uint64_t db = p.gemm ? umma_desc_mn(b_addr) : umma_desc(b_addr);
          const uint32_t idesc_k = p.gemm ? (idesc | (1u << 16)) : idesc;            // gemm: B is MN-major
          const uint64_t bstep = p.gemm ? 128u : 2u;                                 // 16 k-rows = 2 KiB (MN-major) / 32 bytes (K-major), >> 4
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // advance 16 elements = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
            if (PAIR) umma_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)k * bstep, idesc_k, (kb != kb_lo || k != 0) ? 1u : 0u, issue);
            else umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)k * bstep, idesc_k, (kb != kb_lo || k != 0) ? 1u : 0u, issue);
          }
          if (PAIR) umma_commit_pair(&empty[stage], issue); else umma_commit(&empty[stage], issue);   // smem stage free (in both CTAs) once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        for (int j = 0; j < p.res_kb; ++j) {             // D[:, 64j .. 64j+63] += R_j x I
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc(smem_u32(sA + (size_t)stage * TC_A_STAGE));
          const uint64_t db = umma_desc(smem_u32(sEye));
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            if (PAIR) umma_pair(d_tmem + (uint32_t)(j * 64), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_eye, 1u, issue);
            else umma_bf16(d_tmem + (uint32_t)(j * 64), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_eye, 1u, issue);
          }
          if (PAIR) umma_commit_pair(&empty[stage], issue); else umma_commit(&empty[stage], issue);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma_commit_pair(&tmem_full[acc], issue); else umma_commit(&tmem_full[acc], issue);   // accumulator complete (each CTA holds its 128 rows)
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================= epilogue: 8 independent warps =================
    // Warp (grp, quad) owns rows quad*32..+31 (its TMEM lane quadrant) of the 32-column chunks
    // c == grp (mod 2) of every tile (alt_tiles: all chunks of every other tile), with private smem staging, mbarriers and bulk-async groups:
    // TMEM -> registers -> (+bias, +residual tile prefetched by TMA, ReLU, halo) -> smem -> TMA store,
    // no block-level synchronisation anywhere in the epilogue.
    const int grp = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int wslot = warp - 2;                                      // 0..7: private staging buffers / barriers per warp
    const bool elected = lane == 0;
    griddep_wait();
    const bool has_res = p.residual != nullptr && p.res_kb == 0;
    const int chunks_per_tile = (p.BN + 31) / 32;
    const int cstep = p.alt_tiles ? 1 : 2, cfirst = p.alt_tiles ? 0 : grp;   // this warp's chunks: cfirst, cfirst + cstep, ...
    const int my_chunks = p.alt_tiles ? chunks_per_tile : (chunks_per_tile - grp + 1) / 2;
    const int nres = p.nres;                                         // residual buffers per warp (power of two)
    const int nres_shift = nres == 4 ? 2 : (nres == 2 ? 1 : 0);
    uint8_t* gOut = sOut + wslot * (TC_OUT_BUFS * TC_WARP_TILE);
    uint8_t* gRes = sRes + wslot * (nres * TC_WARP_TILE);
    uint64_t* gres_full = res_full + wslot * 4;
    int acc = 0; uint32_t acc_phase = 0;
    int seq = 0;                                                     // warp-local chunk sequence over all tiles
    // residual prefetch cursor: next (tile iteration, local chunk) to request
    int pf_it = 0, pf_j = 0, pf_seq = 0, pf_mt = 0, pf_nt = 0;
    bool pf_ok = p.tma_epi && has_res && my_chunks > 0 && tile_at<PAIR>(p, 0, pf_mt, pf_nt);
    auto issue_res_load = [&]() {
      const int buf = pf_seq & (nres - 1);
      mbar_expect_tx(&gres_full[buf], TC_WARP_TILE);
      tma_load_2d(gRes + buf * TC_WARP_TILE, &tmRes, pf_nt * p.BN + (2 * pf_j + grp) * 32, pf_mt * TC_BM + quad * 32, &gres_full[buf]);
      ++pf_seq;
      if (++pf_j == my_chunks) { pf_j = 0; ++pf_it; pf_ok = tile_at<PAIR>(p, pf_it, pf_mt, pf_nt); }
    };
    if (elected) {
      for (int q = 0; q < nres && pf_ok; ++q) issue_res_load();
    }
    int m_tile, n_tile;
    for (int it = 0; tile_at<PAIR>(p, it, m_tile, n_tile); ++it) {
      if (p.alt_tiles && (it & 1) != grp) {                        // the other group's tile (accumulator acc == it & 1)
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
        continue;
      }
      const long long m = (long long)m_tile * TC_BM + row_in_tile;
      const bool valid = m < p.M;
      int img = 0, y = 0, x = 0;
      bool halo = false;
      if (valid && p.out_mode != 2) {
        const int plane = p.g.plane();
        img = (int)(m / plane);
        const int pos = (int)(m - (long long)img * plane);
        y = pos / p.g.Wp(); x = pos - y * p.g.Wp();
        halo = y == 0 || y == p.g.H + 1 || x == 0 || x == p.g.W + 1;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.BN);
      const int n0 = n_tile * p.BN;
      if (p.tma_epi) {
        const uint32_t sw = (uint32_t)((lane >> 1) & 3);             // SWIZZLE_64B: 16B chunk index ^= address bits [7,8]
        const bool zero_row = halo || !valid;
        const int row0 = m_tile * TC_BM + quad * 32;
        uint32_t r[32];
        if (my_chunks > 0) tmem_ld32(t_base + cfirst * 32, r);         // software pipeline: chunk j+1's TMEM load overlaps chunk j
        for (int j = 0; j < my_chunks; ++j, ++seq) {
          const int c = cstep * j + cfirst;
          const int col = n0 + c * 32;
          const int obuf = seq & (TC_OUT_BUFS - 1);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
            v[i] = b.x; v[i + 1] = b.y; v[i + 2] = b.z; v[i + 3] = b.w;
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(r[i]);
          if (j + 1 < my_chunks) tmem_ld32(t_base + (c + cstep) * 32, r);
          if (has_res) {
            const int rbuf = seq & (nres - 1);
            mbar_wait(&gres_full[rbuf], (uint32_t)((seq >> nres_shift) & 1));
            const uint8_t* rrow = gRes + rbuf * TC_WARP_TILE + lane * 64;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint4 rv = *reinterpret_cast<const uint4*>(rrow + (((uint32_t)jj ^ sw) << 4));
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = unpack2<F16>(rw[q]);
                v[jj * 8 + 2 * q] += f.x; v[jj * 8 + 2 * q + 1] += f.y;
              }
            }
          }
          uint32_t pk[16];
          if (p.relu == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack2_sat<F16, true>(v[2 * i], v[2 * i + 1]);
          } else {
            if (p.relu == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = apply_act(v[i], 2);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack2_sat<F16, false>(v[2 * i], v[2 * i + 1]);
          }
          if (zero_row) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          }
          // gOut[obuf] was last read by this warp's store of chunk seq-2: retire it before overwriting
          if (elected) bulk_wait_read<1>();
          __syncwarp();
          uint8_t* orow = gOut + obuf * TC_WARP_TILE + lane * 64;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            *reinterpret_cast<uint4*>(orow + (((uint32_t)jj ^ sw) << 4)) = make_uint4(pk[jj * 4], pk[jj * 4 + 1], pk[jj * 4 + 2], pk[jj * 4 + 3]);
          fence_async_smem();
          __syncwarp();
          if (elected) {
            tma_store_2d(&tmOut, gOut + obuf * TC_WARP_TILE, col, row0);
            bulk_commit();
            if (pf_ok) issue_res_load();                         // the buffer of chunk seq was consumed before the __syncwarp
          }
        }
      } else {
        // direct-store path (dense fp32 outputs / odd tile widths): same chunk assignment as above
        for (int c0 = cfirst * 32; c0 < p.BN; c0 += 32 * cstep) {
          uint32_t r[32];
          if (c0 + 32 <= p.BN) {
            tmem_ld32(t_base + c0, r);
            tmem_ld_wait();
            epilogue_chunk<32, F16>(p, r, m, n0 + c0, valid, halo, img, y, x);
          } else {
            tmem_ld16(t_base + c0, r);
            tmem_ld_wait();
            epilogue_chunk<16, F16>(p, r, m, n0 + c0, valid, halo, img, y, x);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(map_to_rank(smem_u32(&tmem_empty[acc]), 0)); else mbar_arrive(&tmem_empty[acc]);
      }
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
    if (p.tma_epi && elected) bulk_wait<0>();                     // all output tiles landed before the CTA exits
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();               // pair: neither CTA leaves while the other still uses its smem / TMEM / barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}


// =================================================================================================================
// Fused bottleneck tail + next head (layers.cuh BneckArgs): per 256-row pair tile
//   for each 128-channel chunk c of the expanded width:   A(c):  accA[c&1]  = t2 * W3_c^T  (+ x_c through identity MMAs)
//                                                         E1(c): x'_c = relu(accA + b3_c) -> 16 bit -> shared memory (128B-swizzled
//                                                                k-blocks) -> TMA store to xo AND A operand of
//                                                         B(c):  accB += x'_c * W1_c^T
//   E2: t1 = relu(accB + b1) -> shared memory -> TMA store.
// The MMA warp issues  A(0) A(1) B(0) A(2) B(1) ... A(7) B(6) B(7)  so the tensor pipe works on A(c+1) while the epilogue
// warps turn chunk c around.  Shared memory per CTA: t2 tile resident (kbA x 16 KB), x' staging 2 x 32 KB, a 5-slot ring of
// 16 KB (W3 k-block pairs, residual k-blocks, W1 k-blocks, streamed in exactly the order the MMA warp consumes them), identity
// tile, biases.
// TMEM: 2 x 128 columns accA + 256 columns accB = all 512.  Same pair conventions as k_conv_tc<., true>: loads land in
// the issuing CTA's shared memory and are counted on the LEADER's barriers, commits are multicast to both CTAs, epilogue warps of
// both CTAs arrive on the leader's barriers.
// The residual x rides the ring as two more k-blocks per chunk and is added by the tensor core against an identity tile (the
// res_kb form of k_conv_tc), so the arithmetic -- operand rounding points, k order, residual after the main k-blocks -- is that of
// the two separate k_conv_tc launches and the fused result is BIT-IDENTICAL to the unfused one.  Reading the residual in the epilogue
// instead (one 128-byte row segment per lane and chunk, straight from global memory, double-buffered in registers + L2 prefetch) frees
// a third of the ring and 12 % of the MMA time but was 20-50 % slower: 8 lane-divergent 16-byte loads per thread and chunk through
// an L1 that the 220 KB of shared memory leave 28 KB of (profiles/r2_bneck_experiments.txt).
// =================================================================================================================
constexpr int BK_MAX_SLOTS = 8;
constexpr uint32_t BK_SLOT = 16384;

struct BnParams {
  long long M;
  int m_tiles;              // pair tiles of 256 rows
  int Cmid, Cexp, nchunks, kbA;   // kbA: k-blocks of GEMM A = (Cmid + Cd) / 64
  int kbT;                  // of which from t2 (Cmid / 64); the rest from xd (the block's input, see has_res)
  int has_res;              // 1: residual x added through identity MMAs;  0: the residual branch is a 1x1 convolution of xd folded INTO GEMM A
                            // (first block of a stage: W3 and the downsample weights side by side along K, biases summed)
  int slots;                // ring depth (5..8 slots of 16 KB, whatever shared memory is left)
  int t2_bufs;              // 2: the next tile's t2 is loaded while this one is in use (narrow layers are HBM-latency bound otherwise)
  unsigned long long* trace;  // tooling (YOLACT_B200_BNECK_TRACE): cycle stamps, see k_bneck_tc
  Geom g;
  const float* b3;
  const float* b1;
};

template <bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_bneck_tc(const __grid_constant__ CUtensorMap tmT2, const __grid_constant__ CUtensorMap tmXd, const __grid_constant__ CUtensorMap tmX,
           const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmXo,
           const __grid_constant__ CUtensorMap tmT1, const BnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t rank = cluster_ctarank();
  uint8_t* sT2 = smem;                                         // [t2_bufs][kbA][128 rows x 128 B]
  uint8_t* sX = sT2 + (size_t)p.t2_bufs * p.kbA * TC_A_STAGE;  // [2 buffers][2 k-blocks][128 rows x 128 B]
  uint8_t* sRing = sX + 4 * TC_A_STAGE;                        // [slots][16 KB]
  uint8_t* sEye = sRing + (size_t)p.slots * BK_SLOT;          // [32 rows][128 B]: this CTA's half of the 64 x 64 identity
  float* sBias = reinterpret_cast<float*>(sEye + 4096);        // [Cexp] b3 then [Cmid] b1: read by every epilogue warp, every chunk
  uint64_t* full = reinterpret_cast<uint64_t*>(sBias + p.Cexp + p.Cmid);   // [8]
  uint64_t* empty = full + 8;                                  // [8]
  uint64_t* t2_full = empty + 8;                               // [2]
  uint64_t* t2_empty = t2_full + 2;                            // [2]
  uint64_t* accA_full = t2_empty + 2;                          // [2]
  uint64_t* accA_empty = accA_full + 2;                        // [2]
  uint64_t* accB_full = accA_empty + 2;
  uint64_t* accB_empty = accB_full + 1;
  uint64_t* xs_full = accB_empty + 1;                          // [2]  x' chunk staged by all 16 epilogue warps of the pair
  uint64_t* xs_empty = xs_full + 2;                            // [2]  B(c) finished reading the staging buffer
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xs_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmT2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t2_full[i], 1); mbar_init(&t2_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&accA_full[i], 1); mbar_init(&accA_empty[i], 16);
      mbar_init(&xs_full[i], 16); mbar_init(&xs_empty[i], 1);
    }
    mbar_init(accB_full, 1); mbar_init(accB_empty, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    // identity rows 32*rank .. 32*rank+31 (see k_conv_tc): row n has 1.0 at element 32*rank + n, chunk (k/8) ^ (n & 7)
    const int t = (int)threadIdx.x - 64, n = t >> 2;
    if (n < 32) {
      const int kone = n + 32 * (int)rank;
      const uint32_t one = F16 ? 0x3C00u : 0x3F80u;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = (t & 3) * 2 + q;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (c == (kone >> 3)) {
          const int e = kone & 7;
          const uint32_t w = (e & 1) ? (one << 16) : one;
          if ((e >> 1) == 0) v.x = w; else if ((e >> 1) == 1) v.y = w; else if ((e >> 1) == 2) v.z = w; else v.w = w;
        }
        *reinterpret_cast<uint4*>(sEye + n * 128 + ((c ^ (n & 7)) << 4)) = v;
      }
    }
    fence_async_smem();
    for (int i = t; i < p.Cexp + p.Cmid; i += TC_THREADS - 64) sBias[i] = i < p.Cexp ? __ldg(p.b3 + i) : __ldg(p.b1 + i - p.Cexp);   // constants: before griddep_wait
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();

  const int unit = (int)(blockIdx.x >> 1), units = (int)(gridDim.x >> 1);
  // tooling: cycle stamps of the leader CTA of pair 0 (YOLACT_B200_BNECK_TRACE), 8 events x 64 chunks
  unsigned long long* const trace = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
  auto stamp = [&](int ev, uint32_t qq) { if (trace && qq < 64u) trace[ev * 64 + qq] = (unsigned long long)clock64(); };
  const int nch = p.nchunks, kbA = p.kbA, slots = p.slots;
  const int kpb = kbA >= 2 ? 2 : 1;                                   // W3 k-blocks packed into one ring slot
  const uint32_t w1_bytes = (uint32_t)(p.Cmid / 2) * 128u;            // one W1 k-block of this CTA: Cmid/2 rows x 128 B

  if (warp == 0) {
    // ================= TMA producer (both CTAs; bytes counted on the leader's barriers) =================
    if (lane == 0) {
      griddep_wait();
      int stage = 0; uint32_t phase = 0;
      auto load_a = [&](int c, int row0) {                              // operands of A(c): W3 chunk (kpb k-blocks per slot), residual chunk
        for (int s = 0; s < kbA / kpb; ++s) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t fa = map_to_rank(smem_u32(&full[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full[stage], 2u * (uint32_t)kpb * 8192u);
          for (int j = 0; j < kpb; ++j)
            tma_load_2d_pair(sRing + (size_t)stage * BK_SLOT + (size_t)j * 8192, &tmW3, (kpb * s + j) * TC_BK, c * 128 + (int)rank * 64, fa);
          if (++stage == slots) { stage = 0; phase ^= 1; }
        }
        for (int j = 0; j < 2 && p.has_res; ++j) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t fa = map_to_rank(smem_u32(&full[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * BK_SLOT);
          tma_load_2d_pair(sRing + (size_t)stage * BK_SLOT, &tmX, c * 128 + j * TC_BK, row0, fa);
          if (++stage == slots) { stage = 0; phase ^= 1; }
        }
      };
      auto load_b = [&](int c) {                                        // B operand of B(c): W1[:, c*128 .. +127], two k-blocks
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t fa = map_to_rank(smem_u32(&full[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * w1_bytes);
          tma_load_2d_pair(sRing + (size_t)stage * BK_SLOT, &tmW1, c * 128 + kb * TC_BK, (int)rank * (p.Cmid / 2), fa);
          if (++stage == slots) { stage = 0; phase ^= 1; }
        }
      };
      for (int i = 0; unit + i * units < p.m_tiles; ++i) {
        const int row0 = (2 * (unit + i * units) + (int)rank) * TC_BM;
        const int tb = i % p.t2_bufs;
        const uint32_t tn = (uint32_t)(i / p.t2_bufs);                   // n-th use of this t2 buffer
        mbar_wait(&t2_empty[tb], (tn & 1u) ^ 1u);                        // A(last) of the tile that used it before has read it
        if (rank == 0) mbar_expect_tx(&t2_full[tb], 2u * (uint32_t)kbA * TC_A_STAGE);
        const uint32_t t2_full_addr = map_to_rank(smem_u32(&t2_full[tb]), 0);
        for (int kb = 0; kb < kbA; ++kb)                                 // A tile of GEMM A: t2's k-blocks, then the block input's (if folded in)
          tma_load_2d_pair(sT2 + ((size_t)tb * kbA + kb) * TC_A_STAGE, kb < p.kbT ? &tmT2 : &tmXd, (kb < p.kbT ? kb : kb - p.kbT) * TC_BK, row0, t2_full_addr);
        for (int c = 0; c < nch; ++c) {
          load_a(c, row0);
          if (c >= 1) load_b(c - 1);
        }
        load_b(nch - 1);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA) =================
    if (rank == 0) {
      const uint32_t issue = elect_one();
      const uint32_t fmt = F16 ? 0u : 1u;
      const uint32_t idesc0 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
      const uint32_t idescA = idesc0 | ((uint32_t)(128 >> 3) << 17);
      const uint32_t idescE = idesc0 | ((uint32_t)(64 >> 3) << 17);
      const uint32_t idescB = idesc0 | ((uint32_t)(p.Cmid >> 3) << 17);
      const uint32_t dB = tmem_base + 256u;
      int stage = 0; uint32_t phase = 0;
      uint32_t q = 0;                                                   // chunk sequence number over all tiles of this pair
      auto issue_b = [&](int c, uint32_t qq, int i) {
        const int buf = (int)(qq & 1u);
        if (lane == 0) stamp(2, qq);
        mbar_wait(&xs_full[buf], (qq >> 1) & 1u);                       // x' chunk staged in both CTAs (generic writes + proxy fence)
        if (lane == 0) stamp(3, qq);
        if (c == 0) { mbar_wait(accB_empty, (uint32_t)(i & 1) ^ 1u); }  // E2 of the previous tile has drained accB
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc(smem_u32(sX + (size_t)buf * 2 * TC_A_STAGE + (size_t)kb * TC_A_STAGE));
          const uint64_t db = umma_desc(smem_u32(sRing + (size_t)stage * BK_SLOT));
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_pair(dB, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idescB, (c | kb | k) != 0 ? 1u : 0u, issue);
          umma_commit_pair(&empty[stage], issue);
          if (++stage == slots) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&xs_empty[buf], issue);
      };
      for (int i = 0; unit + i * units < p.m_tiles; ++i) {
        const int tb = i % p.t2_bufs;
        mbar_wait(&t2_full[tb], (uint32_t)(i / p.t2_bufs) & 1u);
        tc_fence_after();
        const uint8_t* sT2b = sT2 + (size_t)tb * kbA * TC_A_STAGE;
        for (int c = 0; c < nch; ++c, ++q) {
          const int a = (int)(q & 1u);
          if (lane == 0) stamp(0, q);
          mbar_wait(&accA_empty[a], ((q >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t dA = tmem_base + (uint32_t)(a * 128);
          for (int s = 0; s < kbA / kpb; ++s) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            for (int j = 0; j < kpb; ++j) {
              const int kb = kpb * s + j;
              const uint64_t da = umma_desc(smem_u32(sT2b + (size_t)kb * TC_A_STAGE));
              const uint64_t db = umma_desc(smem_u32(sRing + (size_t)stage * BK_SLOT + (size_t)j * 8192));
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) umma_pair(dA, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idescA, (kb | k) != 0 ? 1u : 0u, issue);
            }
            umma_commit_pair(&empty[stage], issue);
            if (++stage == slots) { stage = 0; phase ^= 1; }
          }
          for (int j = 0; j < 2 && p.has_res; ++j) {                    // accA[:, 64j .. 64j+63] += x_j * I
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t da = umma_desc(smem_u32(sRing + (size_t)stage * BK_SLOT));
            const uint64_t db = umma_desc(smem_u32(sEye));
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_pair(dA + (uint32_t)(j * 64), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idescE, 1u, issue);
            umma_commit_pair(&empty[stage], issue);
            if (++stage == slots) { stage = 0; phase ^= 1; }
          }
          umma_commit_pair(&accA_full[a], issue);
          if (lane == 0) stamp(1, q);
          if (c == nch - 1) umma_commit_pair(&t2_empty[tb], issue);
          if (c >= 1) issue_b(c - 1, q - 1, i);
        }
        issue_b(nch - 1, q - 1, i);
        umma_commit_pair(accB_full, issue);
      }
    }
  } else {
    // ================= epilogue: warp (grp, quad) owns rows quad*32.. of the 64-column half `grp` of every chunk =================
    const int wslot = warp - 2, grp = wslot >> 2, quad = warp & 3;
    const bool elected = lane == 0;
    griddep_wait();
    uint8_t* const slab0 = sX + (size_t)grp * TC_A_STAGE + (size_t)quad * 4096;      // this warp's 32 x 128 B slab in staging buffer 0; buffer 1 is 32 KB on
    const uint32_t accA_empty_l = map_to_rank(smem_u32(&accA_empty[0]), 0), xs_full_l = map_to_rank(smem_u32(&xs_full[0]), 0);
    const uint32_t accB_empty_l = map_to_rank(smem_u32(accB_empty), 0);
    const uint32_t sw = (uint32_t)(lane & 7);                          // SWIZZLE_128B: 16-byte chunk index ^= row & 7
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t q = 0;
    // one 32-row x 32-column piece: TMEM -> +bias -> ReLU -> 16-bit -> 4 swizzled 16-byte stores into the warp's slab
    auto piece = [&](const uint32_t* r, const float* bias, bool zero_row, uint8_t* row, int h) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(bias + i);
        pk[i / 2] = pack2_sat<F16, true>(__uint_as_float(r[i]) + b.x, __uint_as_float(r[i + 1]) + b.y);
        pk[i / 2 + 1] = pack2_sat<F16, true>(__uint_as_float(r[i + 2]) + b.z, __uint_as_float(r[i + 3]) + b.w);
      }
      if (zero_row) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        *reinterpret_cast<uint4*>(row + ((((uint32_t)(h * 4 + jj)) ^ sw) << 4)) = make_uint4(pk[jj * 4], pk[jj * 4 + 1], pk[jj * 4 + 2], pk[jj * 4 + 3]);
    };
    for (int i = 0; unit + i * units < p.m_tiles; ++i) {
      const int row0 = (2 * (unit + i * units) + (int)rank) * TC_BM + quad * 32;
      const long long m = (long long)row0 + lane;
      bool zero_row = m >= p.M;
      if (!zero_row) {
        const int pos = (int)(m % p.g.plane());
        const int y = pos / p.g.Wp(), x = pos - y * p.g.Wp();
        zero_row = y == 0 || y == p.g.H + 1 || x == 0 || x == p.g.W + 1;
      }
      for (int c = 0; c < nch; ++c, ++q) {
        const int a = (int)(q & 1u);                                    // accumulator buffer == staging buffer of chunk q
        mbar_wait(&accA_full[a], (q >> 1) & 1u);
        if (warp == 2 && lane == 0) stamp(4, q);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t t_addr = t_lane + (uint32_t)(a * 128 + grp * 64);
        tmem_ld32(t_addr, r0);
        tmem_ld32(t_addr + 32, r1);
        mbar_wait(&xs_empty[a], ((q >> 1) & 1u) ^ 1u);                  // B(q-2) has read this staging buffer
        if (elected) {                                                  // ... and so has this warp's TMA store of chunk q-2
          if (c == 0) bulk_wait_read<0>(); else bulk_wait_read<1>();    //     (first chunk of a tile: E2 may have used this slab last)
        }
        __syncwarp();
        tmem_ld_wait();
        if (warp == 2 && lane == 0) stamp(5, q);
        tc_fence_before();
        uint8_t* const slab = slab0 + (size_t)a * 2 * TC_A_STAGE;
        uint8_t* row = slab + lane * 128;
        const float* bias = sBias + c * 128 + grp * 64;
        piece(r0, bias, zero_row, row, 0);
        piece(r1, bias + 32, zero_row, row, 1);
        fence_async_smem();
        __syncwarp();
        if (elected) {
          mbar_arrive_cluster(accA_empty_l + (uint32_t)a * 8u);
          mbar_arrive_cluster(xs_full_l + (uint32_t)a * 8u);
          tma_store_2d(&tmXo, slab, c * 128 + grp * 64, row0);
          bulk_commit();
          if (warp == 2) stamp(6, q);
        }
      }
      // ---- E2: t1 = relu(accB + b1), this warp's Cmid/2 columns: 64-column slabs (staging buffers 0 and 1 in turn), or -- Cmid = 64 --
      //      one 32-column piece in the 64-byte-swizzled form of k_conv_tc's epilogue ----
      mbar_wait(accB_full, (uint32_t)(i & 1));
      tc_fence_after();
      const int cw = p.Cmid / 2;
      if (cw >= 64) {
        for (int j = 0; j < cw / 64; ++j) {
          uint32_t r0[32], r1[32];
          const int col = grp * cw + j * 64;
          tmem_ld32(t_lane + 256u + (uint32_t)col, r0);
          tmem_ld32(t_lane + 256u + (uint32_t)col + 32u, r1);
          if (elected) bulk_wait_read<1>();
          __syncwarp();
          tmem_ld_wait();
          uint8_t* const slab = slab0 + (size_t)(j & 1) * 2 * TC_A_STAGE;
          uint8_t* row = slab + lane * 128;
          piece(r0, sBias + p.Cexp + col, zero_row, row, 0);
          piece(r1, sBias + p.Cexp + col + 32, zero_row, row, 1);
          fence_async_smem();
          __syncwarp();
          if (elected) {
            tma_store_2d(&tmT1, slab, col, row0);
            bulk_commit();
          }
        }
      } else {
        uint32_t r0[32];
        const int col = grp * 32;
        tmem_ld32(t_lane + 256u + (uint32_t)col, r0);
        if (elected) bulk_wait_read<1>();
        __syncwarp();
        tmem_ld_wait();
        const float* bias = sBias + p.Cexp + col;
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + e);
          pk[e / 2] = pack2_sat<F16, true>(__uint_as_float(r0[e]) + bb.x, __uint_as_float(r0[e + 1]) + bb.y);
          pk[e / 2 + 1] = pack2_sat<F16, true>(__uint_as_float(r0[e + 2]) + bb.z, __uint_as_float(r0[e + 3]) + bb.w);
        }
        if (zero_row) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = 0u;
        }
        const uint32_t sw64 = (uint32_t)((lane >> 1) & 3);             // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,8]
        uint8_t* row = slab0 + lane * 64;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          *reinterpret_cast<uint4*>(row + (((uint32_t)jj ^ sw64) << 4)) = make_uint4(pk[jj * 4], pk[jj * 4 + 1], pk[jj * 4 + 2], pk[jj * 4 + 3]);
        fence_async_smem();
        __syncwarp();
        if (elected) {
          tma_store_2d(&tmT1, slab0, col, row0);
          bulk_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (elected) mbar_arrive_cluster(accB_empty_l);
    }
    if (elected) bulk_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

static int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint32_t box_rows, bool f16,
                    uint32_t box_inner = TC_BK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, uint64_t row_stride = 0) {
  EncodeTiledFn enc = get_encode();
  YB_REQUIRE(enc != nullptr, YB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {inner, rows};
  const cuuint64_t strides[1] = {(row_stride ? row_stride : inner) * 2};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YB_REQUIRE(r == CUDA_SUCCESS, YB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%llu rows=%llu box_rows=%u", (int)r,
             (unsigned long long)inner, (unsigned long long)rows, box_rows);
  return YB_OK;
}

bool tc_overlapping_rows_ok() {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  alignas(64) static CUtensorMap probe;
  const cuuint64_t dims[2] = {64, 4096};
  const cuuint64_t strides[1] = {32};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  void* base = nullptr;
  if (cudaMalloc(&base, 4096 * 32 + 128) != cudaSuccess) { cudaGetLastError(); return false; }
  const CUresult r = enc(&probe, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFree(base);
  return r == CUDA_SUCCESS;
}

static int tc_device_setup(int* sms);

static int pick_bn(int cout_pad) {
  if (const char* e = getenv("YOLACT_B200_BN")) { const int v = atoi(e); if (v >= 32 && v <= 256 && v % 32 == 0 && cout_pad % v == 0) return v; }   // tooling
  if (cout_pad % 256 == 0) return 256;
  for (int bn = 224; bn >= 64; bn -= 32)                               // largest 32-multiple divisor: 1152 -> 192, 288 -> 96
    if (cout_pad % bn == 0 && cout_pad > 256) return bn;
  if (cout_pad <= 256 && cout_pad % 16 == 0) return cout_pad;          // 32, 64, 128, ...
  if (cout_pad % 2 == 0 && (cout_pad / 2) % 16 == 0 && cout_pad / 2 <= 256) return cout_pad / 2;   // 352 -> 176
  return 0;
}

bool tc_supported(const ConvArgs& a) {
  if (a.act_dt != DT_BF16 && a.act_dt != DT_F16) return false;
  if (a.Cin % 8 != 0 || a.Cin_pad % TC_BK != 0 || a.Cin_pad < a.Cin) return false;
  if (a.out_mode == 0 && a.Cout != a.Cout_pad) return false;
  return pick_bn(a.Cout_pad) != 0;
}

int tc_plan_create(const ConvArgs& a, int max_batch, TcPlan** out) {
  YB_REQUIRE(tc_supported(a), YB_ERR_UNSUPPORTED, "tc_plan_create: unsupported conv Cin=%d Cout_pad=%d", a.Cin, a.Cout_pad);
  TcPlan* pl = new TcPlan();
  pl->BN = pick_bn(a.Cout_pad);
  int cols = 32;
  while (cols < 2 * pl->BN) cols <<= 1;
  pl->tmem_cols = cols;
  pl->tma_epi = (a.out_mode == 0 && pl->BN % 32 == 0) ? 1 : 0;
  // CTA pair (cluster of 2, tcgen05 cta_group::2): one 256 x BN tile per pair, each CTA stages its own 128 A rows and HALF of the
  // weight tile -- a third less L2 -> shared-memory fill per MMA than two independent CTAs, and a half-size resident weight
  // slice.  Measured per layer (profiles/r1_pair_vs_single.txt, r1_conv_tc_experiments.txt): -5..-12 % where the layer is
  // fill-bound (K >= 1024 with the full 256-wide N tile: the 3x3 and the 1024/2048-channel reduce convs) and -5..-19 % on the
  // 256-wide expand convs with a residual (resident half slice + deep slab ring); neutral or slightly worse elsewhere (the pair
  // advances at the pace of its slower CTA).  YOLACT_B200_PAIR=0 / 1 forces it off / on for every eligible (BN % 32 == 0) layer.
  const int ktot = a.ntaps * a.Cin_pad;
  pl->pair = (pl->BN == 256 && (ktot >= 1024 || (a.residual && a.out_mode == 0 && ktot >= 256))) ? 1 : 0;
  if (const char* e = getenv("YOLACT_B200_PAIR")) pl->pair = (atoi(e) != 0 && pl->BN % 32 == 0) ? 1 : 0;
  const size_t b_stage = (size_t)(pl->pair ? pl->BN / 2 : pl->BN) * TC_BK * 2;
  const int num_kb = a.ntaps * a.Cin_pad / TC_BK;
  const size_t budget = 227 * 1024 - 1024 /*align*/ - 1024 /*barriers*/;
  // residual through the tensor core: measured in the network it pays for the short-K expand convs (64->256 @138: -8 %, 128->512 @69:
  // -18 %) whose epilogue is the bottleneck, and costs 4-20 % once K >= 256, where the extra A-slab stages compete with the main
  // k-blocks for the few pipeline stages.  YOLACT_B200_RESMMA=0 / 1 forces it off / on.
  // In the pair form the half weight slice of a K = 256 expand (64 KB) stays resident next to 7 slab stages, so there the residual
  // slabs ride the pipeline for free: 256->1024 @35 + residual 87 -> 73 us.
  bool res_mma = ktot <= 128 || (pl->pair && ktot <= 256);
  if (const char* e = getenv("YOLACT_B200_RESMMA")) res_mma = atoi(e) != 0;
  pl->res_kb = (pl->tma_epi && a.residual && pl->BN % 64 == 0 && res_mma) ? pl->BN / 64 : 0;
  pl->nres = (pl->tma_epi && a.residual && !pl->res_kb) ? 2 : 0;   // residual buffers per epilogue group
  if (pl->nres) if (const char* e = getenv("YOLACT_B200_NRES")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) pl->nres = v; }
  size_t epi_bytes = pl->tma_epi ? (size_t)8 * (TC_OUT_BUFS + pl->nres) * TC_WARP_TILE : 0;
  if (pl->res_kb) epi_bytes += 8192;                               // identity operand tile
  pl->alt_tiles = (pl->BN <= 64 && pl->nres == 0 && !getenv("YOLACT_B200_NO_ALT")) ? 1 : 0;
  // stem (net.cu: 4 tap rows over the space-to-depth image, rows overlapping at 16 of 64 channels): the 32-byte-row form
  pl->stem16 = (a.in_row_stride == 16 && a.Cin == 64 && a.Cin_pad == 64 && a.ntaps == 4 && pl->BN == 64 && !pl->pair && !a.residual &&
                !getenv("YOLACT_B200_NO_STEM16")) ? 1 : 0;
  // slab mode (3x3, stride 1: the three dx taps of a tap row are consecutive rows of the same matrix)
  pl->slab = (a.ntaps == 9 && !pl->res_kb && !getenv("YOLACT_B200_NO_SLAB")) ? 1 : 0;
  if (pl->stem16) pl->slab = 0;
  for (int dy = 0; dy < 3 && pl->slab; ++dy)
    for (int dx = 1; dx < 3; ++dx)
      if (a.tap_shift[3 * dy + dx] != a.tap_shift[3 * dy] + dx) pl->slab = 0;
  pl->stages_a = 0;
  const size_t a_slot = pl->slab ? TC_A_SLAB : TC_A_STAGE;
  // weight-resident mode: the whole [BN x Ktot] slice fits next to >= 3 A stages
  pl->b_resident = 0;
  const int n_tiles = a.Cout_pad / pl->BN;
  if (!getenv("YOLACT_B200_NO_BRES") && (size_t)num_kb * b_stage + epi_bytes + 3 * a_slot <= budget && n_tiles <= 64) {
    pl->b_resident = 1;
    int stages = (int)((budget - epi_bytes - (size_t)num_kb * b_stage) / a_slot);
    stages = stages > 8 ? 8 : stages;
    if (pl->slab) { pl->stages_a = stages; pl->stages = 1; } else pl->stages = stages;
    pl->smem_bytes = (size_t)stages * a_slot + (size_t)num_kb * b_stage + epi_bytes + 1024 + 1024;
  } else if (pl->slab) {
    // separate rings: 3 slabs (= 9 k-blocks of look-ahead on the A side), the rest of the budget for weight tiles
    pl->stages_a = 3;
    int stages = (int)((budget - epi_bytes - 3 * a_slot) / b_stage);
    if (stages > 8) { stages = 8; pl->stages_a = (int)((budget - epi_bytes - 8 * b_stage) / a_slot); if (pl->stages_a > 8) pl->stages_a = 8; }
    pl->stages = stages;
    pl->smem_bytes = (size_t)pl->stages_a * a_slot + (size_t)stages * b_stage + epi_bytes + 1024 + 1024;
    if (stages < 3) { pl->slab = 0; pl->stages_a = 0; }
  }
  if (pl->stem16) {                                                // 16 resident weight tiles of 2 KB, 8 slab slots of 5 KB
    pl->b_resident = 1; pl->stages = 1; pl->stages_a = 8;
    pl->smem_bytes = (size_t)8 * TC_S16_SLOT + 16 * 2048 + epi_bytes + 1024 + 1024;
  }
  if (!pl->b_resident && !pl->slab) {
    const size_t per_stage = TC_A_STAGE + b_stage;
    int stages = (int)((budget - epi_bytes) / per_stage);
    pl->stages = stages > 8 ? 8 : stages;
    pl->smem_bytes = (size_t)pl->stages * per_stage + epi_bytes + 1024 + 1024;
  }
  const int Ktot = a.ntaps * a.Cin_pad;
  const int cout_alloc = (a.Cout_pad + 63) / 64 * 64;
  const bool f16 = a.act_dt == DT_F16;
  int s = pl->stem16 ? make_map(&pl->tmA, a.in, 16, (uint64_t)a.in_rows + 4, TC_S16_ROWS, f16, 16, CU_TENSOR_MAP_SWIZZLE_32B)
                     : make_map(&pl->tmA, a.in, (uint64_t)a.Cin, (uint64_t)a.in_rows, pl->slab ? TC_SLAB_ROWS : TC_BM, f16, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B, (uint64_t)a.in_row_stride);
  if (s == YB_OK && pl->stem16) s = make_map(&pl->tmB, a.weight, (uint64_t)Ktot, (uint64_t)cout_alloc, 64, f16, 16, CU_TENSOR_MAP_SWIZZLE_32B, (uint64_t)a.w_ld);
  else if (s == YB_OK) s = make_map(&pl->tmB, a.weight, (uint64_t)Ktot, (uint64_t)cout_alloc, (uint32_t)(pl->pair ? pl->BN / 2 : pl->BN), f16, TC_BK,
                               CU_TENSOR_MAP_SWIZZLE_128B, (uint64_t)a.w_ld);
  pl->tmOut = pl->tmA; pl->tmRes = pl->tmA;                      // placeholders when the TMA epilogue is off
  if (s == YB_OK && pl->tma_epi) {
    const uint64_t out_rows = (uint64_t)max_batch * a.g.plane();
    s = make_map(&pl->tmOut, a.out, (uint64_t)a.Cout, out_rows, 32, f16, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (s == YB_OK && a.residual) {
      if (pl->res_kb) s = make_map(&pl->tmRes, a.residual, (uint64_t)a.Cout, out_rows, TC_BM, f16);   // A-operand slabs
      else s = make_map(&pl->tmRes, a.residual, (uint64_t)a.Cout, out_rows, 32, f16, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    }
  }
  if (s != YB_OK) { delete pl; return s; }
  { const int st = tc_device_setup(&pl->sms); if (st != YB_OK) { delete pl; return st; } }
  pl->gemm = 0;
  *out = pl;
  return YB_OK;
}

void tc_plan_destroy(TcPlan* p) { delete p; }

// per-device one-time setup shared by the conv and GEMM plans: 227 KB of dynamic shared memory for every instantiation
static int tc_device_setup(int* sms) {
  int dev = 0;
  YB_CHECK_CUDA(cudaGetDevice(&dev));
  static std::mutex mu;
  static bool done[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  if (dev >= 0 && dev < 64 && !done[dev]) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    YB_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    YB_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    YB_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    YB_CHECK_CUDA(cudaFuncSetAttribute(k_bneck_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    YB_CHECK_CUDA(cudaFuncSetAttribute(k_bneck_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    done[dev] = true;
  }
  YB_CHECK_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  YB_REQUIRE(*sms > 0, YB_ERR_CUDA, "cannot read the SM count");
  return YB_OK;
}

// ---- plain GEMM over long K on the same kernel (weight gradients; see TcParams::gemm) ---------------------------
int tc_plan_create_gemm(const GemmArgs& g, TcPlan** out) {
  YB_REQUIRE(g.act_dt == DT_BF16 || g.act_dt == DT_F16, YB_ERR_UNSUPPORTED, "tc_plan_create_gemm: 16-bit operands only");
  YB_REQUIRE(g.M >= 1 && g.K >= 1 && g.lda % 8 == 0 && g.ldb % 8 == 0 && g.ldb >= g.Nper && g.ntaps >= 1 && g.ntaps <= 16, YB_ERR_INVALID,
             "tc_plan_create_gemm: M=%d K=%d lda=%d ldb=%d ntaps=%d", g.M, g.K, g.lda, g.ldb, g.ntaps);
  int bn = 0;
  if (g.Nper % 256 == 0) bn = 256;
  else if (g.Nper % 128 == 0) bn = 128;
  else if (g.Nper % 64 == 0) bn = 64;
  YB_REQUIRE(bn != 0, YB_ERR_UNSUPPORTED, "tc_plan_create_gemm: N per tap = %d is not a multiple of 64", g.Nper);
  TcPlan* pl = new TcPlan();
  pl->gemm = 1; pl->BN = bn;
  int cols = 32;
  while (cols < 2 * bn) cols <<= 1;
  pl->tmem_cols = cols;
  pl->tma_epi = 0; pl->nres = 0; pl->b_resident = 0; pl->res_kb = 0; pl->pair = 0; pl->slab = 0; pl->stages_a = 0;
  pl->alt_tiles = bn <= 64 ? 1 : 0;
  const size_t b_stage = (size_t)bn * TC_BK * 2, per_stage = TC_A_STAGE + b_stage;
  const size_t budget = 227 * 1024 - 2048;
  int stages = (int)(budget / per_stage);
  pl->stages = stages > 8 ? 8 : stages;
  pl->smem_bytes = (size_t)pl->stages * per_stage + 2048;
  const bool f16 = g.act_dt == DT_F16;
  int s = make_map(&pl->tmA, g.a, (uint64_t)g.K, (uint64_t)g.M, TC_BM, f16, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B, (uint64_t)g.lda);
  // B is read MN-major straight from the [Kb pixels][ldb channels] activations: boxes of [64 k-rows][64 channels]
  if (s == YB_OK) s = make_map(&pl->tmB, g.b, (uint64_t)g.Nper, (uint64_t)g.Kb, 64, f16, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B, (uint64_t)g.ldb);
  pl->tmOut = pl->tmA; pl->tmRes = pl->tmA;
  if (s == YB_OK) s = tc_device_setup(&pl->sms);
  if (s != YB_OK) { delete pl; return s; }
  *out = pl;
  return YB_OK;
}

int launch_gemm_tc(const TcPlan* pl, const GemmArgs& g, cudaStream_t s) {
  YB_REQUIRE(pl && pl->gemm, YB_ERR_INVALID, "launch_gemm_tc: not a GEMM plan");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = g.M;
  p.m_tiles = (g.M + TC_BM - 1) / TC_BM;
  p.gemm = 1; p.gemm_ntile_tap = g.Nper / pl->BN;
  p.gemm_splits = g.splits > 0 ? g.splits : 1;
  p.n_tiles = g.ntaps * p.gemm_ntile_tap;
  p.kb_per_tap = (g.K + TC_BK - 1) / TC_BK;
  p.gemm_kb_split = (p.kb_per_tap + p.gemm_splits - 1) / p.gemm_splits;
  p.gemm_splits = (p.kb_per_tap + p.gemm_kb_split - 1) / p.gemm_kb_split;          // no empty split
  YB_REQUIRE(p.gemm_splits == 1 || g.accumulate, YB_ERR_INVALID, "launch_gemm_tc: split-K needs accumulate (atomic) output");
  p.ntaps = 1; p.tap_shift[0] = 0;
  for (int i = 0; i < g.ntaps; ++i) p.gemm_shift[i] = g.shift[i];
  p.BN = pl->BN; p.tmem_cols = pl->tmem_cols; p.stages = pl->stages;
  p.Cout = p.Cout_pad = g.ntaps * g.Nper; p.relu = 0; p.out_mode = 2; p.accumulate = g.accumulate;
  p.is_f16 = g.act_dt == DT_F16; p.alt_tiles = pl->alt_tiles;
  p.g.H = 1 << 20; p.g.W = 1 << 20;
  p.bias = nullptr; p.residual = nullptr; p.out = g.out;
  const int total = p.m_tiles * p.n_tiles * p.gemm_splits;
  const int grid = total < pl->sms ? total : pl->sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = pl->smem_bytes; cfg.stream = s;
  cfg.attrs = nullptr; cfg.numAttrs = 0;
  const cudaError_t le = p.is_f16 ? cudaLaunchKernelEx(&cfg, k_conv_tc<true, false>, pl->tmA, pl->tmB, pl->tmOut, pl->tmRes, p)
                                  : cudaLaunchKernelEx(&cfg, k_conv_tc<false, false>, pl->tmA, pl->tmB, pl->tmOut, pl->tmRes, p);
  YB_CHECK_CUDA(le);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

int launch_conv_tc(const TcPlan* pl, const ConvArgs& a, cudaStream_t s) {
  TcParams p;
  p.M = (long long)a.B * a.g.plane();
  const int unit_rows = pl->pair ? 2 * TC_BM : TC_BM;
  p.m_tiles = (int)((p.M + unit_rows - 1) / unit_rows);            // tiles per scheduling unit (CTA or CTA pair)
  p.n_tiles = a.Cout_pad / pl->BN;
  p.kb_per_tap = a.Cin_pad / TC_BK;
  p.ntaps = a.ntaps;
  for (int i = 0; i < kMaxTaps; ++i) p.tap_shift[i] = a.tap_shift[i];
  p.BN = pl->BN; p.tmem_cols = pl->tmem_cols; p.stages = pl->stages;
  p.Cout = a.Cout; p.Cout_pad = a.Cout_pad; p.relu = a.relu; p.out_mode = a.out_mode;
  p.g = a.g; p.bias = a.bias; p.residual = a.residual; p.out = a.out; p.is_f16 = a.act_dt == DT_F16; p.tma_epi = pl->tma_epi; p.nres = pl->nres; p.b_resident = pl->b_resident; p.res_kb = pl->res_kb; p.slab = pl->slab; p.stages_a = pl->stages_a; p.alt_tiles = pl->alt_tiles; p.stem16 = pl->stem16;
  p.gemm = 0; p.gemm_ntile_tap = 1; p.accumulate = 0; p.gemm_splits = 1; p.gemm_kb_split = 0;
  for (int i = 0; i < 16; ++i) p.gemm_shift[i] = 0;
  const int sms = pl->sms;
  const int total = p.m_tiles * p.n_tiles;
  const int max_units = pl->pair ? sms / 2 : sms;
  int units = total < max_units ? total : max_units;
  if (pl->b_resident) {
    units = max_units / p.n_tiles * p.n_tiles;                // whole groups of N tiles
    if (units > total) units = total;                         // total is a multiple of n_tiles
    if (units < p.n_tiles) units = p.n_tiles;
  }
  const int grid = pl->pair ? 2 * units : units;
  static const bool pdl = getenv("YOLACT_B200_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = pl->smem_bytes; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (pl->pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t le;
  if (pl->pair) le = p.is_f16 ? cudaLaunchKernelEx(&cfg, k_conv_tc<true, true>, pl->tmA, pl->tmB, pl->tmOut, pl->tmRes, p)
                              : cudaLaunchKernelEx(&cfg, k_conv_tc<false, true>, pl->tmA, pl->tmB, pl->tmOut, pl->tmRes, p);
  else le = p.is_f16 ? cudaLaunchKernelEx(&cfg, k_conv_tc<true, false>, pl->tmA, pl->tmB, pl->tmOut, pl->tmRes, p)
                     : cudaLaunchKernelEx(&cfg, k_conv_tc<false, false>, pl->tmA, pl->tmB, pl->tmOut, pl->tmRes, p);
  YB_CHECK_CUDA(le);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

// ---- fused bottleneck tail (k_bneck_tc) ------------------------------------------------------------------------------
struct BnPlan {
  CUtensorMap tmT2, tmXd, tmX, tmW3, tmW1, tmXo, tmT1;
  int sms, Cmid, Cexp, Cd, slots, t2_bufs;
  size_t smem_bytes;
};

bool bneck_supported(int act_dt, int Cmid, int Cexp) {
  // Cmid <= 256: accB (Cmid columns) sits next to the two 128-column accA buffers in the 512 TMEM columns; Cexp / 128 chunks must be
  // even (accumulator / staging buffers alternate by chunk parity across tiles)
  return (act_dt == DT_F16 || act_dt == DT_BF16) && (Cmid == 64 || Cmid == 128 || Cmid == 256) && Cexp == 4 * Cmid && !getenv("YOLACT_B200_NO_FUSE");
}

// shared-memory plan: [t2 buffers][x' staging 64 KB][ring][identity 4 KB][biases][barriers]
static void bneck_smem(int Cmid, int Cexp, int Cd, int* slots, int* t2_bufs, size_t* bytes) {
  const size_t t2 = (size_t)((Cmid + Cd) / 64) * TC_A_STAGE;
  const size_t fixed = 1024 + 4 * TC_A_STAGE + 4096 + (size_t)(Cexp + Cmid) * 4 + 1024;
  const size_t budget = 227 * 1024;
  *t2_bufs = (fixed + 2 * t2 + 5 * BK_SLOT <= budget) ? 2 : 1;
  int n = (int)((budget - fixed - (size_t)*t2_bufs * t2) / BK_SLOT);
  *slots = n > BK_MAX_SLOTS ? BK_MAX_SLOTS : n;
  *bytes = fixed + (size_t)*t2_bufs * t2 + (size_t)*slots * BK_SLOT;
}

int bneck_plan_create(const BneckArgs& a, int max_batch, BnPlan** out) {
  YB_REQUIRE(bneck_supported(a.act_dt, a.Cmid, a.Cexp), YB_ERR_UNSUPPORTED, "bneck_plan_create: unsupported Cmid=%d Cexp=%d", a.Cmid, a.Cexp);
  BnPlan* pl = new BnPlan();
  pl->Cmid = a.Cmid; pl->Cexp = a.Cexp; pl->Cd = a.Cd;
  YB_REQUIRE(a.Cd == 0 || (a.xd && a.Cd % 64 == 0 && (a.Cmid + a.Cd) / 64 <= 2), YB_ERR_UNSUPPORTED, "bneck_plan_create: folded residual branch with Cd=%d", a.Cd);
  YB_REQUIRE(a.Cd != 0 || a.x, YB_ERR_INVALID, "bneck_plan_create: no residual");
  const bool f16 = a.act_dt == DT_F16;
  const uint64_t rows = (uint64_t)max_batch * a.g.plane();
  int s = make_map(&pl->tmT2, a.t2, (uint64_t)a.Cmid, rows, TC_BM, f16);
  pl->tmXd = pl->tmT2; pl->tmX = pl->tmT2;                        // placeholders for the operand a form does not have
  if (s == YB_OK && a.Cd) s = make_map(&pl->tmXd, a.xd, (uint64_t)a.Cd, rows, TC_BM, f16);
  if (s == YB_OK && !a.Cd) s = make_map(&pl->tmX, a.x, (uint64_t)a.Cexp, rows, TC_BM, f16);
  if (s == YB_OK) s = make_map(&pl->tmW3, a.w3, (uint64_t)(a.Cmid + a.Cd), (uint64_t)a.Cexp, 64, f16);
  if (s == YB_OK) s = make_map(&pl->tmW1, a.w1, (uint64_t)a.Cexp, (uint64_t)a.Cmid, (uint32_t)(a.Cmid / 2), f16);
  if (s == YB_OK) s = make_map(&pl->tmXo, a.xo, (uint64_t)a.Cexp, rows, 32, f16);
  if (s == YB_OK) s = a.Cmid >= 128 ? make_map(&pl->tmT1, a.t1, (uint64_t)a.Cmid, rows, 32, f16)
                                    : make_map(&pl->tmT1, a.t1, (uint64_t)a.Cmid, rows, 32, f16, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (s == YB_OK) s = tc_device_setup(&pl->sms);
  if (s != YB_OK) { delete pl; return s; }
  bneck_smem(a.Cmid, a.Cexp, a.Cd, &pl->slots, &pl->t2_bufs, &pl->smem_bytes);
  YB_REQUIRE(pl->slots >= 5 && pl->smem_bytes <= 227 * 1024, YB_ERR_UNSUPPORTED, "bneck_plan_create: %zu bytes of shared memory, %d slots", pl->smem_bytes, pl->slots);
  *out = pl;
  return YB_OK;
}

void bneck_plan_destroy(BnPlan* p) { delete p; }

static unsigned long long* g_bneck_trace = nullptr;     // tooling: stamps of the most recent launch (yb_debug_bneck_trace)

int launch_bneck_tc(const BnPlan* pl, const BneckArgs& a, cudaStream_t s) {
  BnParams p;
  p.M = (long long)a.B * a.g.plane();
  p.m_tiles = (int)((p.M + 2 * TC_BM - 1) / (2 * TC_BM));
  p.Cmid = a.Cmid; p.Cexp = a.Cexp; p.nchunks = a.Cexp / 128; p.kbT = a.Cmid / 64; p.kbA = (a.Cmid + pl->Cd) / 64; p.has_res = pl->Cd ? 0 : 1;
  p.slots = pl->slots; p.t2_bufs = pl->t2_bufs;
  p.g = a.g; p.b3 = a.b3; p.b1 = a.b1;
  p.trace = nullptr;
  if (getenv("YOLACT_B200_BNECK_TRACE")) {
    if (!g_bneck_trace) { YB_CHECK_CUDA(cudaMalloc(&g_bneck_trace, 8 * 64 * 8)); }
    YB_CHECK_CUDA(cudaMemsetAsync(g_bneck_trace, 0, 8 * 64 * 8, s));
    p.trace = g_bneck_trace;
  }
  const int max_units = pl->sms / 2;
  const int units = p.m_tiles < max_units ? p.m_tiles : max_units;
  static const bool pdl = getenv("YOLACT_B200_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * units)); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = pl->smem_bytes; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attr; cfg.numAttrs = na;
  const cudaError_t le = a.act_dt == DT_F16
      ? cudaLaunchKernelEx(&cfg, k_bneck_tc<true>, pl->tmT2, pl->tmXd, pl->tmX, pl->tmW3, pl->tmW1, pl->tmXo, pl->tmT1, p)
      : cudaLaunchKernelEx(&cfg, k_bneck_tc<false>, pl->tmT2, pl->tmXd, pl->tmX, pl->tmW3, pl->tmW1, pl->tmXo, pl->tmT1, p);
  YB_CHECK_CUDA(le);
  YB_CHECK_LAUNCH();
  return YB_OK;
}

}  // namespace yb

// tooling (tools/bneck_trace.py): copy out the cycle stamps of the most recent k_bneck_tc launch made with YOLACT_B200_BNECK_TRACE set
extern "C" __attribute__((visibility("default"))) int yb_debug_bneck_trace(unsigned long long* out, int count) {
  if (!yb::g_bneck_trace || count > 8 * 64) return -1;
  return cudaMemcpy(out, yb::g_bneck_trace, (size_t)count * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
