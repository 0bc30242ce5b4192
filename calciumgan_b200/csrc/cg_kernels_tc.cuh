// tcgen05 / TMEM / TMA implicit-GEMM kernels for the bf16 mixed-precision path (sm_100a).
//
//  rsgemm_tc : out[b,q,n] = epi( sum_{tap,c} A[b, q+shift(tap), acol(tap)+c] * W[n, wk(tap)+c] )
//              strided Conv1D fwd, Conv1DTranspose fwd, their data gradients, GP linearised fwd, Dense.
//              A tiles are tap-shifted 3-D TMA boxes (batch is its own dim -> per-sample zero fill),
//              W tiles 2-D TMA boxes, both K-major SWIZZLE_128B; D (128 x BN fp32) lives in TMEM,
//              double buffered; warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue (two per TMEM lane quarter).
//  rsgemm3_tc: the same GEMM on CTA pairs (cta_group::2, M = 256), slab reuse through row-shifted descriptors, row-pair and
//              merged-phase forms for 64-channel layers; serves every layer with >= 128 time rows per sample.
//  wgrad_tc  : dW[tap][m][n] += sum_{b,q} S[b, q+shift(tap), scol(tap)+m] * P[b,q,n]
//              both operands MN-major (reduction runs over time rows), split over rows.
//  wgrad2_tc : slab reuse (one slab per 64-row chunk serves all taps of a parity group through LBO-shifted descriptors);
//              epilogue = accumulator rows staged in the idle stage ring + one bulk reduce-add per row segment.
//  wgrad2p_tc: wgrad2 on CTA pairs: two work items share the P tile, each CTA loads / reads half of it (default).
#pragma once
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <utility>

#include "cg_common.cuh"

// Role cycle counters (CG_TC_TIMING=1 at run time) need the clock reads compiled in: build with -DCG_TC_INSTRUMENT
// (CG_TC_INSTRUMENT=1 calciumgan_b200/csrc/build.sh). The shipped build has none: ~350 clock reads per CTA sat on the
// critical path of the single MMA-issuing thread.
#ifdef CG_TC_INSTRUMENT
#define CG_CLK() clock64()
#define CG_DBG_ON (P.dbg != nullptr)
#else
#define CG_CLK() 0LL
#define CG_DBG_ON false
#endif

#ifdef CG_TC_INSTRUMENT
#define CG_TC_INSTRUMENTED 1
#else
#define CG_TC_INSTRUMENTED 0
#endif

#ifndef CG_TC_SPIN_LIMIT
#define CG_TC_SPIN_LIMIT (1u << 26)   // bounded mbarrier spin: trap instead of hanging the GPU
#endif

namespace tc {

// ---------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > CG_TC_SPIN_LIMIT) {
      printf("calciumgan_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// descriptor passed as 32-bit halves (lo = start address | LBO field, hi = SBO | version | swizzle): the issue
// loop only ever adds to `lo`, which keeps the per-MMA instruction count minimal
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                               uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %5, 0;\n mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accum)
      : "memory");
}
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start (and
// run its prologue: barrier init, TMEM allocation, tensor-map prefetch) while its predecessor in the stream drains;
// griddep_wait() blocks until the predecessor has completed and its writes are visible (no-op without the attribute).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// one elected lane of a fully converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);   // version 1 (bit 46), SWIZZLE_128B (bits 61-63)
}
// ---- CTA-pair (cta_group::2) helpers --------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local smem address` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in the executing CTA's smem, bytes are counted on the barrier at `mbar_cluster`
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint32_t mbar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(mbar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint32_t mbar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(mbar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA over the CTA pair (issued by the leader CTA only)
__device__ __forceinline__ void umma2_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                                uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %5, 0;\n mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n"
      " tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accum)
      : "memory");
}
// completion of all prior MMAs of this thread -> arrive on the barrier at the same smem offset in both CTAs
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((unsigned short)3)
               : "memory");
}

// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp reads TMEM lane (32*(warp%4) + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (sm_100). Offsets in bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// instruction descriptor for kind::f16: bf16 x bf16 -> fp32, M x N, a/b major (0 = K-major, 1 = MN-major)
__host__ __device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

constexpr int kThreads = 320;               // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kEpiWarps = 8;
constexpr int kWgThreads = 192;             // weight-gradient kernel v1: warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kWg2Threads = 320;            // wgrad2: eight epilogue warps (two per TMEM lane quarter, alternate 32-column chunks):
                                            // the red.add tail of a CTA, during which its tensor pipe idles, was 12-15 % of the kernel
constexpr int kABytes = 128 * 128;          // 128 rows x 64 bf16
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;             // columns between the two accumulator buffers

// explicit shared-space accesses for the epilogue staging buffer: with generic pointers the compiler must assume
// the staging stores alias the global loads/stores around them and serialises every global round-trip
__device__ __forceinline__ void sts_v4(uint32_t sa, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ uint4 lds_v4(uint32_t sa) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t sa, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(sa), "f"(v)); }
__device__ __forceinline__ float lds_f32(uint32_t sa) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sa)); return v; }
// 16-byte global -> shared copy without registers; src_bytes = 0 writes zeros
__device__ __forceinline__ void cp_async16(uint32_t sa, const void* g, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5 with the hardware tanh approximation: ONE MUFU op per element instead of
// EX2 + RCP (abs error ~1e-4, far inside the bf16 path's 2e-2; the fp32 path uses expf)
__device__ __forceinline__ float sigmoid_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

// ---- epilogue -----------------------------------------------------------------------------------------
// Eight epilogue warps: warp w reads TMEM lanes 32 * (w % 4) .. +31 (hardware rule), and the two warps of a lane
// quarter take alternate 32-column chunks (half = 0 / 1). One warp per scheduler left every TMEM / shared / global
// latency of the epilogue exposed (measured 2.4k - 4.9k clk per 128 x 64 block against 3k - 6k clk of MMA work).
constexpr int kStgBytes = 2048;   // per-warp staging: 32 rows x 32 bf16 (64-byte rows), 16-byte chunks XOR-swizzled by (row >> 1) & 3
constexpr int kEpiSmem = kEpiWarps * kStgBytes + 3 * 1024;   // + bias / gamma / beta tiles (256 floats each)

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // all 8 epilogue warps
__device__ __forceinline__ void epi_bar_half(int half) {                                       // the 4 warps of one column half
  asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
}
__device__ __forceinline__ uint32_t stg_addr(uint32_t stg_s, int row, int chunk) {
  return stg_s + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}

// Per tile, each lane precomputes the element offsets of the 4 rows it stores in the coalesced phase
// (row rr = 8*i + lane/4 of this warp's 32 rows); negative = masked. rpt < 128 is a power of two.
struct EpiRows {
  long long base;                     // element offset of sample b0; off[] / my_off are 32-bit offsets relative to it
  int off[4]; int my_off; long long my_o32; long long my_row; bool my_ok;
  long long o32_base; bool uniform;   // single-sample blocks: fp32 row rr of this warp is at o32_base + rr * o32_rs
  int x_src, x_par; bool x_zero;      // EPI_PS_MASK: exchange slot this row fills / adds (-1 = none), row without a source
};
__device__ __forceinline__ void epi_rows(const RsParams& p, int b0, int q0, int rpt_log2, int phase, int lq, int lane,
                                         EpiRows& R) {
  const int crow = lane >> 2;
  const int ph_off = phase * p.o_phase_col;
  R.base = (long long)b0 * p.o_bs;
#pragma unroll
  for (int i = 0; i <= 4; ++i) {
    const int r = lq * 32 + (i < 4 ? i * 8 + crow : lane);
    int db, q;
    if (rpt_log2 >= 7) { db = 0; q = q0 + r; } else { db = r >> rpt_log2; q = r & ((1 << rpt_log2) - 1); }
    const bool ok = b0 + db < p.B;
    const int o = ok ? db * (int)p.o_bs + q * p.o_rs + ph_off : -1;   // per-sample extents are far below 2^31 elements
    if (i < 4) R.off[i] = o;
    else {
      const int b = b0 + db;
      R.my_off = o; R.my_ok = ok; R.my_o32 = (long long)b * p.o32_bs + (long long)q * p.o32_rs;
      R.my_row = ((long long)b * p.Q + q) * p.seg.nphase + phase;
    }
  }
  R.uniform = rpt_log2 >= 7;
  R.o32_base = (long long)b0 * p.o32_bs + (long long)(q0 + lq * 32) * p.o32_rs;
  R.x_src = -1; R.x_par = -1; R.x_zero = false;
}

// EPI_PS_MASK rows (single-sample 128-row blocks of one output phase): accumulator row t = (q0 + r) * nphase + phase goes
// to time j = t + s when that is inside [0, w); rows pushed over an edge are reflected (calciumgan.py:126-133) onto a
// row of the SAME block and phase (t + t' is even and both lie within 2|s| <= 20 steps of the edge), so the sum is formed
// in registers through a small shared-memory exchange: the reflected row fills slot x_src, its partner adds slot x_par.
// Rows nobody maps to (x_zero) are written as zeros by the thread that owns the same index.
__device__ __forceinline__ void epi_rows_ps(const RsParams& p, int b, int q0, int phase, int s, int lq, int lane, EpiRows& R) {
  const int w = p.Q * 2;   // two output phases (host-checked): time t = 2 q + phase
  const int crow = lane >> 2;
  const bool bok = b < p.B;
  R.base = (long long)b * p.o_bs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int j, xs, xp; bool xz;
    ps_adjoint_row((q0 + lq * 32 + i * 8 + crow) * 2 + phase, s, w, j, xs, xp, xz);
    R.off[i] = (bok && j >= 0) ? (j >> 1) * p.o_rs + (j & 1) * p.o_phase_col : -1;
  }
  const int q = q0 + lq * 32 + lane;
  const int t = q * 2 + phase;
  R.my_ok = bok;
  R.my_off = q * p.o_rs + phase * p.o_phase_col;
  R.my_o32 = 0; R.my_row = 0; R.o32_base = 0; R.uniform = true;
  int dest;
  ps_adjoint_row(t, s, w, dest, R.x_src, R.x_par, R.x_zero);
}

// The 32-column chunks c0 = 32 * half, 32 * half + 64, ... of one 128-row x BN-column accumulator block:
// TMEM -> registers (thread = row) -> bias / LeakyReLU / sigmoid / layer-norm / slope mask in fp32 -> bf16 -> swizzled
// smem transpose -> global stores where every warp instruction writes eight 64-byte row segments. EPI is a
// compile-time mode: the first version of this epilogue spent ~10k clk per 128x64 block on per-element mode checks,
// integer divisions and 64-bit address math (profiles/). Padded columns come out as exact zeros because padded
// weight rows and the staged bias are zero.
template <int EPI>
__device__ __forceinline__ void epilogue_block(const RsParams& p, uint32_t tmem_cols, int BN, int n_base,
                                               EpiRows& R, const float* bias_s, uint8_t* stg, uint8_t* stg_partner,
                                               int lq, int lane, int half, int ps_q0 = 0, int ps_s = 0, bool need_x = false,
                                               int ps_b = 0) {
  // ps_q0: time index of tile row 0 (single-sample blocks); ps_s: this sample's PhaseShuffle shift
  bf16* out = reinterpret_cast<bf16*>(p.out);
  bf16* psx = reinterpret_cast<bf16*>(p.ps_out);
  const uint32_t stg_s = smem_u32(stg);
  const bf16* mask = reinterpret_cast<const bf16*>(p.mask);
  const int cj = lane & 3, crow = lane >> 2;
  const uint32_t my_st = stg_s + lane * 64;
  const int msw = (lane >> 1) & 3;
  const uint32_t taddr = tmem_cols + ((uint32_t)(lq * 32) << 16);
  float ssq = 0.f;   // p.sumsq: gradient-penalty norm fused into the last data-gradient GEMM (fp32 accumulators)
  float ln_mean = 0.f, ln_rstd = 1.f;
  if (EPI == EPI_BIAS_LN_LRELU) {   // pass 1 over TMEM: row statistics of x = acc + bias; this thread sees half of its row
    float s1 = 0.f, s2 = 0.f;
    for (int c0 = half * 32; c0 < BN; c0 += 64) {
      uint32_t v[32];
      tmem_ld32(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (n_base + c0 + j < p.n_real) {
          const float x = __uint_as_float(v[j]) + bias_s[c0 + j];
          s1 += x;
          s2 = fmaf(x, x, s2);
        }
      }
    }
    // the other half of the row lives in the partner warp (same lane quarter): swap partial sums through the staging buffers
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stg_s + lane * 8), "f"(s1), "f"(s2) : "memory");
    epi_bar();
    float o1, o2;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(o1), "=f"(o2) : "r"(smem_u32(stg_partner) + lane * 8) : "memory");
    epi_bar();   // everybody has read before the staging buffers are reused
    s1 += o1; s2 += o2;
    ln_mean = s1 / p.n_real;
    ln_rstd = rsqrtf(fmaxf(s2 / p.n_real - ln_mean * ln_mean, 0.f) + CG_LN_EPS);
    if (half == 0 && p.mu && R.my_ok) { p.mu[R.my_row] = ln_mean; p.rstd[R.my_row] = ln_rstd; }
  }
  for (int c0 = half * 32; c0 < BN; c0 += 64) {
    int n0 = n_base + c0;
    if (EPI == EPI_PS_MASK && p.merged_phases) {   // column half = output phase: this chunk's rows follow that phase's map
      epi_rows_ps(p, ps_b, ps_q0, n0 >> 6, ps_s, lq, lane, R);
      n0 &= 63;
    }
    if (EPI == EPI_MASK || EPI == EPI_PS_MASK) {   // coalesced, register-free read of the slope source (same indexing as out) into the staging tile
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = i * 8 + crow;
        const bool ok = R.off[i] >= 0;
        cp_async16(stg_addr(stg_s, rr, cj), mask + (ok ? R.base + R.off[i] + n0 + cj * 8 : 0), ok ? 16u : 0u);
      }
    }
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    tmem_ld_wait();
    if (c0 + 32 > BN) {   // BN is a multiple of 16, not of 32 (e.g. 112 for 102 channels): columns past the tile are not ours
#pragma unroll
      for (int j = 16; j < 32; ++j) v[j] = 0u;
    }
    if (EPI == EPI_PS_MASK && need_x) {   // reflected rows meet their partners (bias tile area: 2 halves x 2 x 5 slots x 32 floats)
      const uint32_t xb = smem_u32(bias_s) + half * 1280 + ((c0 >> 6) & 1) * 640;
      if (R.x_src >= 0) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) sts_v4(xb + R.x_src * 128 + j4 * 16, make_uint4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]));
      }
      epi_bar_half(half);
      if (R.x_par >= 0) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const uint4 o = lds_v4(xb + R.x_par * 128 + j4 * 16);
          v[j4 * 4] = __float_as_uint(__uint_as_float(v[j4 * 4]) + __uint_as_float(o.x));
          v[j4 * 4 + 1] = __float_as_uint(__uint_as_float(v[j4 * 4 + 1]) + __uint_as_float(o.y));
          v[j4 * 4 + 2] = __float_as_uint(__uint_as_float(v[j4 * 4 + 2]) + __uint_as_float(o.z));
          v[j4 * 4 + 3] = __float_as_uint(__uint_as_float(v[j4 * 4 + 3]) + __uint_as_float(o.w));
        }
      }
    }
    if (EPI == EPI_MASK || EPI == EPI_PS_MASK) {
      cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 mv = lds_v4(my_st + ((j ^ msw) << 4));
        const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
        for (int w2 = 0; w2 < 4; ++w2) {
          const __nv_bfloat162 hv = *reinterpret_cast<const __nv_bfloat162*>(&mw[w2]);
          v[j * 8 + w2 * 2] = __float_as_uint(__uint_as_float(v[j * 8 + w2 * 2]) * lrelu_slope(__low2float(hv)));
          v[j * 8 + w2 * 2 + 1] = __float_as_uint(__uint_as_float(v[j * 8 + w2 * 2 + 1]) * lrelu_slope(__high2float(hv)));
        }
      }
      __syncwarp();
    }
    if (EPI == EPI_BIAS_LN_LRELU) {
      const float* gamma_s = bias_s + 256;
      const float* beta_s = bias_s + 512;
      bf16* aux = reinterpret_cast<bf16*>(p.aux);
      if (aux) {   // pre-norm activations for the backward pass, through the same coalescing transpose
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[j * 8 + e]) + bias_s[c0 + j * 8 + e];
          uint4 o;
          o.x = pack_bf16x2(x[0], x[1]); o.y = pack_bf16x2(x[2], x[3]);
          o.z = pack_bf16x2(x[4], x[5]); o.w = pack_bf16x2(x[6], x[7]);
          sts_v4(my_st + ((j ^ msw) << 4), o);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 o = lds_v4(stg_addr(stg_s, i * 8 + crow, cj));
          if (R.off[i] >= 0) *reinterpret_cast<uint4*>(aux + R.base + R.off[i] + n0 + cj * 8) = o;
        }
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float a = ln_rstd * gamma_s[c0 + j];                        // pads: gamma = beta = 0
        const float d = fmaf(bias_s[c0 + j] - ln_mean, a, beta_s[c0 + j]);
        v[j] = __float_as_uint(lrelu(fmaf(__uint_as_float(v[j]), a, d)));
      }
    }
    if (EPI == EPI_BIAS || EPI == EPI_BIAS_LRELU || EPI == EPI_BIAS_SIGMOID) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + j4 * 4);   // smem broadcast
        const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = __uint_as_float(v[j4 * 4 + e]) + bv[e];
          if (EPI == EPI_BIAS_LRELU) x = lrelu(x);
          if (EPI == EPI_BIAS_SIGMOID) x = (n0 + j4 * 4 + e < p.n_real) ? sigmoid_fast(x) : 0.f;
          v[j4 * 4 + e] = __float_as_uint(x);
        }
      }
    }
    if (EPI == EPI_NONE && p.sumsq) {
#pragma unroll
      for (int j = 0; j < 32; ++j) ssq = fmaf(__uint_as_float(v[j]), __uint_as_float(v[j]), ssq);
    }
    if (p.out32) {   // unpadded fp32 copy (generic generator head): two 32 x 16 fp32 transposes through the staging buffer so
                     // every warp store writes 16 consecutive floats of two output rows
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int j = 0; j < 16; ++j) sts_f32(stg_s + (lane * 16 + (j ^ (lane & 15))) * 4, __uint_as_float(v[h * 16 + j]));
        __syncwarp();
        const int n = n0 + h * 16 + (lane & 15);
        const bool n_ok = n < p.n_real;
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
          const int rr = 2 * it + (lane >> 4);
          const float val = lds_f32(stg_s + (rr * 16 + ((lane & 15) ^ (rr & 15))) * 4);
          if (R.uniform) {   // all 32 rows belong to one sample: plain address arithmetic
            if (R.my_ok && n_ok) p.out32[R.o32_base + (long long)rr * p.o32_rs + n] = val;
          } else {
            const long long o = __shfl_sync(0xffffffffu, R.my_o32, rr);
            const int ok = __shfl_sync(0xffffffffu, (int)R.my_ok, rr);
            if (ok && n_ok) p.out32[o + n] = val;
          }
        }
        __syncwarp();
      }
    }
    if (out || psx) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
        o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
        o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
        o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
        sts_v4(my_st + ((j ^ msw) << 4), o);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = i * 8 + crow;
        const uint4 o = lds_v4(stg_addr(stg_s, rr, cj));
        if (R.off[i] >= 0) {
          if (out) *reinterpret_cast<uint4*>(out + R.base + R.off[i] + n0 + cj * 8) = o;
          if (psx) {   // scatter form of the PhaseShuffle gather: row q feeds every t with ps_index(t) == q
            int t1, t2;
            if (p.row_pairs) {   // GEMM row = time steps (2 qs, 2 qs + 1); this 32-column chunk belongs to one of them
              const int q = 2 * (ps_q0 + lq * 32 + rr) + (n0 >> 6);
              ps_scatter_targets(q, ps_s, p.ps_w, t1, t2);
              bf16* dst = psx + R.base + (n0 & 63) + cj * 8;
              if (t1 >= 0) *reinterpret_cast<uint4*>(dst + t1 * (p.o_rs >> 1)) = o;
              if (t2 >= 0) *reinterpret_cast<uint4*>(dst + t2 * (p.o_rs >> 1)) = o;
            } else {
              const int q = ps_q0 + lq * 32 + rr;
              ps_scatter_targets(q, ps_s, p.ps_w, t1, t2);
              if (t1 >= 0) *reinterpret_cast<uint4*>(psx + R.base + (R.off[i] + (t1 - q) * p.o_rs + n0 + cj * 8)) = o;
              if (t2 >= 0) *reinterpret_cast<uint4*>(psx + R.base + (R.off[i] + (t2 - q) * p.o_rs + n0 + cj * 8)) = o;
            }
          }
        }
      }
      __syncwarp();
    }
    if (EPI == EPI_PS_MASK && R.x_zero && R.my_ok) {   // times no t maps to
#pragma unroll
      for (int c = 0; c < 32; c += 8) *reinterpret_cast<uint4*>(out + R.base + R.my_off + n0 + c) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (EPI == EPI_NONE && p.sumsq) {
    if (R.uniform) {   // all 32 rows of the warp belong to one sample
      ssq = warp_sum(R.my_ok ? ssq : 0.f);
      if (lane == 0 && R.my_ok) atomicAdd(&p.sumsq[R.my_row / ((long long)p.Q * p.seg.nphase)], (double)ssq);
    } else if (R.my_ok) {
      atomicAdd(&p.sumsq[R.my_row / ((long long)p.Q * p.seg.nphase)], (double)ssq);
    }
  }
}

// stage bias[n_base .. n_base+BN) (zero beyond n_real) into smem for the 8 epilogue warps (tid = 0..255)
template <int EPI>
__device__ __forceinline__ void epi_load_bias(const RsParams& p, float* bias_s, int n_base, int BN, int tid) {
  if (EPI == EPI_BIAS || EPI == EPI_BIAS_LRELU || EPI == EPI_BIAS_SIGMOID || EPI == EPI_BIAS_LN_LRELU) {
    epi_bar();   // previous tile's readers are done
    {
      const int n = n_base + tid;
      const int ch = p.row_pairs ? (n & 63) : n;   // row pairs: column = (time parity, channel)
      const bool ok = tid < BN && ch < p.n_real;
      bias_s[tid] = ok ? __ldg(&p.bias[ch]) : 0.f;
      if (EPI == EPI_BIAS_LN_LRELU) {
        bias_s[256 + tid] = ok ? __ldg(&p.gamma[ch]) : 0.f;
        bias_s[512 + tid] = ok ? __ldg(&p.beta[ch]) : 0.f;
      }
    }
    epi_bar();
  }
}

struct RsTcParams {
  RsParams p;
  int BN, n_tiles, m_tiles, rpt, bpt, tiles_per_sample, kchunks, stages;
  int a_bytes, x_shift, x_baseoff;   // x_baseoff: timing-experiment bits (1 no stores, 2 no W loads, 4 no A loads)
  long long* dbg;                    // optional per-role cycle counters of CTA 0 (CG_TC_TIMING=1)
};

// =============================================================================================
// rsgemm_tc
// =============================================================================================
template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
rsgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ RsTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const RsParams& p = P.p;
  const int BN = P.BN;
  const int stages = P.stages;
  const int stage_bytes = P.a_bytes + BN * 128;
  uint8_t* tiles = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* stage_buf = smem + (size_t)stages * stage_bytes + 512;   // 4 x kStgBytes epilogue staging + bias tile
  float* bias_s = reinterpret_cast<float*>(stage_buf + kEpiWarps * kStgBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.seg.nphase * P.n_tiles * P.m_tiles;

  if (warp == 0) {
    int stage = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int mt = t % P.m_tiles;
      const int rest = t / P.m_tiles;
      const int nt = rest % P.n_tiles;
      const int phase = rest / P.n_tiles;
      int b0, q0;
      if (P.tiles_per_sample > 0) { b0 = mt / P.tiles_per_sample; q0 = (mt % P.tiles_per_sample) * 128; }
      else { b0 = mt * P.bpt; q0 = 0; }
      const int nseg = p.seg.nseg[phase];
      for (int s = 0; s < nseg; ++s) {
        const int row = q0 + p.seg.shift[phase][s];
        const int acol = p.seg.acol[phase][s];
        const int wk = p.seg.wk[phase][s];
        for (int kc = 0; kc < P.kchunks; ++kc) {
          const long long tw0 = CG_CLK();
          mbar_wait(&empty[stage], ph ^ 1);
          if (CG_DBG_ON && blockIdx.x == 0 && lane == 0) { atomicAdd((unsigned long long*)&P.dbg[0], (unsigned long long)(CG_CLK() - tw0)); atomicAdd((unsigned long long*)&P.dbg[1], 1ull); }
          if (elect_one()) {
            uint8_t* sa = tiles + (size_t)stage * stage_bytes;
            const uint32_t tx = ((P.x_baseoff & 4) ? 0u : (uint32_t)P.a_bytes) + ((P.x_baseoff & 2) ? 0u : (uint32_t)(BN * 128));
            if (tx) mbar_expect_tx(&full[stage], tx); else mbar_arrive(&full[stage]);
            if (!(P.x_baseoff & 4)) tma_load_3d(sa, &tmA, &full[stage], acol + kc * 64, row, b0);
            if (!(P.x_baseoff & 2)) tma_load_2d(sa + P.a_bytes, &tmW, &full[stage], wk + kc * 64, nt * BN);
          }
          if (++stage == stages) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, BN, 0, 0);
    const uint32_t hi = desc_hi(1024);
    const uint32_t lo0 = desc_lo(smem_u32(tiles), 16);
    const uint32_t stage_step = (uint32_t)stage_bytes >> 4, b_off = (uint32_t)P.a_bytes >> 4;
    int stage = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int phase = (t / P.m_tiles) / P.n_tiles;
      const int acc = it & 1;
      long long tw0 = CG_CLK();
      mbar_wait(&tempty[acc], (((uint32_t)it >> 1) & 1) ^ 1);
      if (CG_DBG_ON && blockIdx.x == 0 && lane == 0) { atomicAdd((unsigned long long*)&P.dbg[2], (unsigned long long)(CG_CLK() - tw0)); atomicAdd((unsigned long long*)&P.dbg[3], 1ull); }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccStride;
      const int kiters = p.seg.nseg[phase] * P.kchunks;
      for (int ki = 0; ki < kiters; ++ki) {
        tw0 = CG_CLK();
        mbar_wait(&full[stage], ph);
        if (CG_DBG_ON && blockIdx.x == 0 && lane == 0) { atomicAdd((unsigned long long*)&P.dbg[4], (unsigned long long)(CG_CLK() - tw0)); atomicAdd((unsigned long long*)&P.dbg[5], 1ull); }
        tc_fence_after();
        const uint32_t a_lo = lo0 + stage * stage_step;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzle row
            umma_bf16_lohi(d_tmem, a_lo + 2 * k, a_lo + b_off + 2 * k, hi, idesc, (uint32_t)(ki | k));
          umma_commit(&empty[stage]);
          if (ki == kiters - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; ph ^= 1; }
      }
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31  == tile rows
    const int lq = warp & 3, half = (warp - 2) >> 2;
    uint8_t* stg = stage_buf + (warp - 2) * kStgBytes;
    uint8_t* stg_partner = stage_buf + ((warp - 2) ^ 4) * kStgBytes;
    int rpt_log2 = 7;
    if (P.tiles_per_sample == 0) { rpt_log2 = 0; while ((1 << rpt_log2) < P.rpt) ++rpt_log2; }
    int it = 0, last_nt = -1;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int mt = t % P.m_tiles;
      const int rest = t / P.m_tiles;
      const int nt = rest % P.n_tiles;
      const int phase = rest / P.n_tiles;
      int b0, q0;
      if (P.tiles_per_sample > 0) { b0 = mt / P.tiles_per_sample; q0 = (mt % P.tiles_per_sample) * 128; }
      else { b0 = mt * P.bpt; q0 = 0; }
      EpiRows R;
      epi_rows(p, b0, q0, rpt_log2, phase, lq, lane, R);
      if (nt != last_nt) { epi_load_bias<EPI>(p, bias_s, nt * BN, BN, threadIdx.x - 64); last_nt = nt; }
      const int acc = it & 1;
      const long long te0 = CG_CLK();
      mbar_wait(&tfull[acc], ((uint32_t)it >> 1) & 1);
      const long long te1 = CG_CLK();
      tc_fence_after();
      epilogue_block<EPI>(p, tmem_base + acc * kAccStride, BN, nt * BN, R, bias_s, stg, stg_partner, lq, lane, half);
      if (CG_DBG_ON && blockIdx.x == 0 && threadIdx.x == 64) {
        atomicAdd((unsigned long long*)&P.dbg[6], (unsigned long long)(te1 - te0));
        atomicAdd((unsigned long long*)&P.dbg[7], (unsigned long long)(CG_CLK() - te1));
        atomicAdd((unsigned long long*)&P.dbg[8], 1ull);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =============================================================================================
// rsgemm2_tc: slab-reuse variant (time rows per sample >= 128).
//  * per (64-channel chunk, tap group) ONE activation slab = MB boxes of (128 + span) rows is loaded; every
//    tap of the group reads its 128-row window through a UMMA descriptor that starts `shift` rows into the
//    slab (SWIZZLE_128B is a function of the absolute smem address, so any 128-byte row start is legal with
//    base_offset = 0 - verified on B200). Activation traffic from L2 drops by the number of taps per group.
//  * MB (1|2) accumulators of 128 rows share each weight tile, halving weight bytes per MAC.
// =============================================================================================
#ifdef CG_TC_INSTRUMENT
// counters accumulate in registers and are flushed once at kernel end: per-stage global atomics from the single
// MMA-issuing thread perturbed the pipeline they were measuring (~900 clk per weight stage)
#define CG_DBG_ADD(i, t0) do { dacc[i] += CG_CLK() - (t0); dacc[(i) + 8] += 1; } while (0)
#define CG_DBG_DECL long long dacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define CG_DBG_FLUSH do { if (CG_DBG_ON && blockIdx.x == 0 && lane == 0) { _Pragma("unroll") for (int q_ = 0; q_ < 16; ++q_) if (q_ != 7 && dacc[q_]) atomicAdd((unsigned long long*)&P.dbg[q_], (unsigned long long)dacc[q_]); } } while (0)
#else
#define CG_DBG_ADD(i, t0) do { (void)(t0); } while (0)
#define CG_DBG_DECL
#define CG_DBG_FLUSH do { } while (0)
#endif
struct SlabGroup {
  int acol, min_shift, nseg;
  unsigned char shift_rel[32];
  int wk[32];
};
struct RsTc2Params {
  RsParams p;
  int BN, n_tiles, m_tiles, MB, blocks_per_sample, total_blocks, kchunks;
  int box_rows, box_bytes, slab_bytes, slab_stages, b_stages, double_acc;
  int acc_stride;       // TMEM columns between the two accumulator buffers (BN rounded up to 32)
  int dbg_nk;           // timing experiment: K steps per chunk (0 = all)
  int dbg_flags;        // timing experiment (slab mode): 1 no epilogue, 2 no weight loads, 4 no activation loads
  int tps;              // taps per weight stage (pair kernel): more MMAs per barrier round-trip for narrow N
  int per_tap;          // 1: rows per sample < 128 -> no slab reuse: a slab stage holds one 128-row box PER TAP (tps boxes)
  int rpt, bpt, rpt_log2;
  int ngroups[2];
  SlabGroup grp[2][4];   // tap groups sharing one slab (same column offset): 2 for the stride-2 forms, 4 for row pairs
  long long* dbg;   // optional role cycle counters of CTA 0 (CG_TC_TIMING=1)
};

// =============================================================================================
// rsgemm3_tc: CTA-pair (cta_group::2) version of rsgemm2. A cluster of two CTAs computes a 256-row x BN tile:
// each CTA owns one 128-row block (its own slab, its own TMEM accumulator and epilogue) and HALF of every weight
// tile; the leader issues tcgen05.mma.cta_group::2 (M = 256), so each SM reads only BN/2 weight rows per MMA and
// receives half the weight bytes from TMA -- the measured limiter of the single-CTA kernel is shared-memory
// bandwidth: (A 4 KB + B N*32 B + TMA fill share) / 128 B/clk per MMA.
// Barrier protocol: full barriers live in the leader (both CTAs' TMA loads complete_tx there); empty / accumulator
// -full barriers are multicast-committed to both CTAs; accumulator-empty is counted on the leader (8 warps).
// =============================================================================================
template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
rsgemm3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ RsTc2Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const RsParams& p = P.p;
  const int BN = P.BN;
  const int SS = P.slab_stages, BS = P.b_stages;
  const int tap_bytes = BN * 64;               // this CTA's half of one tap's weight tile: BN/2 rows x 128 B
  const int TPS = P.tps;
  const int b_bytes = TPS * tap_bytes;         // one weight stage carries TPS taps
  uint8_t* slabs = smem;
  uint8_t* btiles = smem + (size_t)SS * P.slab_bytes;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(btiles + (size_t)BS * b_bytes);
  uint64_t* s_empty = s_full + SS;
  uint64_t* b_full = s_empty + SS;
  uint64_t* b_empty = b_full + BS;
  uint64_t* tfull = b_empty + BS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* stage_buf = btiles + (size_t)BS * b_bytes + 512;
  float* bias_s = reinterpret_cast<float*>(stage_buf + kEpiWarps * kStgBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int i = 0; i < SS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
    for (int i = 0; i < BS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();              // everything above overlapped the predecessor's tail; global memory is touched below
  griddep_launch();

  const int total_tiles = p.seg.nphase * P.n_tiles * P.m_tiles;   // m_tiles = ceil(total_blocks / 2)
  const bool PT = P.per_tap != 0;
  // stage structure: slab mode  -> groups = tap-parity groups, each split into weight stages of TPS taps sharing one slab;
  //                  per-tap mode -> groups = ceil(nseg / TPS) bundles of TPS taps, each tap with its own 16 KB box

  const long long t_kernel0 = CG_CLK();
  CG_DBG_DECL
  if (warp == 0) {
    int ss = 0, bs = 0;
    uint32_t sph = 0, bph = 0;
    for (int t = pair; t < total_tiles; t += npairs) {
      const int mt = t % P.m_tiles;
      const int rest = t / P.m_tiles;
      const int nt = rest % P.n_tiles;
      const int phase = rest / P.n_tiles;
      const int blk = mt * 2 + (int)rank;
      const int b = PT ? blk * P.bpt : blk / P.blocks_per_sample;
      const int q0 = PT ? 0 : (blk % P.blocks_per_sample) * 128;
      for (int kc = 0; kc < P.kchunks; ++kc) {
        if (PT) {
          const int nseg = p.seg.nseg[phase];
          for (int s = 0; s < nseg; s += TPS) {
            const int cnt = nseg - s < TPS ? nseg - s : TPS;
            mbar_wait(&s_empty[ss], sph ^ 1);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(&s_full[ss], (uint32_t)(2 * cnt * kABytes));
              const uint32_t bar = mapa_u32(smem_u32(&s_full[ss]), 0);
              for (int j = 0; j < cnt; ++j)
                tma2_load_3d(slabs + (size_t)ss * P.slab_bytes + (size_t)j * kABytes, &tmA, bar,
                             p.seg.acol[phase][s + j] + kc * 64, q0 + p.seg.shift[phase][s + j], b);
            }
            if (++ss == SS) { ss = 0; sph ^= 1; }
            mbar_wait(&b_empty[bs], bph ^ 1);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(&b_full[bs], (uint32_t)(2 * cnt * tap_bytes));
              const uint32_t bar = mapa_u32(smem_u32(&b_full[bs]), 0);
              for (int j = 0; j < cnt; ++j)
                tma2_load_2d(btiles + (size_t)bs * b_bytes + (size_t)j * tap_bytes, &tmW, bar,
                             p.seg.wk[phase][s + j] + kc * 64, nt * BN + (int)rank * (BN / 2));
            }
            if (++bs == BS) { bs = 0; bph ^= 1; }
          }
          continue;
        }
        for (int g = 0; g < P.ngroups[phase]; ++g) {
          const SlabGroup& G = P.grp[phase][g];
          mbar_wait(&s_empty[ss], sph ^ 1);
          if (elect_one()) {
            if (P.dbg_flags & 4) { if (rank == 0) mbar_arrive(&s_full[ss]); }   // timing experiment: no activation loads
            else {
              if (rank == 0) mbar_expect_tx(&s_full[ss], (uint32_t)(2 * P.slab_bytes));
              tma2_load_3d(slabs + (size_t)ss * P.slab_bytes, &tmA, mapa_u32(smem_u32(&s_full[ss]), 0), G.acol + kc * 64,
                           q0 + G.min_shift, b);
            }
          }
          if (++ss == SS) { ss = 0; sph ^= 1; }
          for (int s = 0; s < G.nseg; s += TPS) {
            const int cnt = G.nseg - s < TPS ? G.nseg - s : TPS;
            mbar_wait(&b_empty[bs], bph ^ 1);
            if (elect_one()) {
              if (P.dbg_flags & 2) { if (rank == 0) mbar_arrive(&b_full[bs]); }   // timing experiment: no weight loads
              else {
                if (rank == 0) mbar_expect_tx(&b_full[bs], (uint32_t)(2 * cnt * tap_bytes));
                const uint32_t bar = mapa_u32(smem_u32(&b_full[bs]), 0);
                for (int j = 0; j < cnt; ++j)
                  tma2_load_2d(btiles + (size_t)bs * b_bytes + (size_t)j * tap_bytes, &tmW, bar, G.wk[s + j] + kc * 64,
                               nt * BN + (int)rank * (BN / 2));
              }
            }
            if (++bs == BS) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc(256, BN, 0, 0);
      const uint32_t hi = desc_hi(1024);
      const uint32_t slab_lo0 = desc_lo(smem_u32(slabs), 16), b_lo0 = desc_lo(smem_u32(btiles), 16);
      const uint32_t slab_step = (uint32_t)P.slab_bytes >> 4, b_step = (uint32_t)b_bytes >> 4;
      const uint32_t tap_step = (uint32_t)tap_bytes >> 4;
      int ss = 0, bs = 0;
      uint32_t sph = 0, bph = 0;
      int it = 0;
      int nk_last = (p.k_real - (P.kchunks - 1) * 64 + 15) >> 4;
      if (nk_last < 1 || nk_last > 4) nk_last = 4;
      for (int t = pair; t < total_tiles; t += npairs, ++it) {
        const int phase = (t / P.m_tiles) / P.n_tiles;
        const int acc = it & 1;
        long long tq = CG_CLK();
        mbar_wait(&tempty[acc], (((uint32_t)it >> 1) & 1) ^ 1);
        CG_DBG_ADD(2, tq);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * P.acc_stride;
        uint32_t accum = 0;
        for (int kc = 0; kc < P.kchunks; ++kc) {
          // K = 16 steps of this 64-channel chunk that hold real channels (the rest multiplies exact zeros: 102 -> 7 of 8)
          const int nk = P.dbg_nk ? P.dbg_nk : (kc == P.kchunks - 1 ? nk_last : 4);
          if (PT) {
            const int nseg = p.seg.nseg[phase];
            for (int s = 0; s < nseg; s += TPS) {
              const int cnt = nseg - s < TPS ? nseg - s : TPS;
              mbar_wait(&s_full[ss], sph);
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint32_t sl_lo = slab_lo0 + ss * slab_step;
              const uint32_t b_lo = b_lo0 + bs * b_step;
              if (elect_one()) {
                for (int j = 0; j < cnt; ++j) {
                  const uint32_t a_lo = sl_lo + j * (kABytes >> 4);
                  const uint32_t bj = b_lo + j * tap_step;
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (k < nk) umma2_bf16_lohi(d_tmem, a_lo + 2 * k, bj + 2 * k, hi, idesc, accum | (uint32_t)(j | k));
                }
                umma2_commit_mc(&b_empty[bs]);
                umma2_commit_mc(&s_empty[ss]);
              }
              __syncwarp();
              accum = 1;
              if (++bs == BS) { bs = 0; bph ^= 1; }
              if (++ss == SS) { ss = 0; sph ^= 1; }
            }
            continue;
          }
          for (int g = 0; g < P.ngroups[phase]; ++g) {
            const SlabGroup& G = P.grp[phase][g];
            tq = CG_CLK();
            mbar_wait(&s_full[ss], sph);
            CG_DBG_ADD(3, tq);
            tc_fence_after();
            const uint32_t sl_lo = slab_lo0 + ss * slab_step;
            for (int s = 0; s < G.nseg; s += TPS) {
              const int cnt = G.nseg - s < TPS ? G.nseg - s : TPS;
              tq = CG_CLK();
              mbar_wait(&b_full[bs], bph);
              CG_DBG_ADD(4, tq);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + bs * b_step;
              if (elect_one()) {
                const long long tm0 = CG_CLK(); (void)tm0;
                for (int j = 0; j < cnt; ++j) {
                  const uint32_t a_lo = sl_lo + (uint32_t)G.shift_rel[s + j] * 8;
                  const uint32_t bj = b_lo + j * tap_step;
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (k < nk) umma2_bf16_lohi(d_tmem, a_lo + 2 * k, bj + 2 * k, hi, idesc, accum | (uint32_t)(j | k));
                }
                const long long tm1 = CG_CLK(); (void)tm1;
                umma2_commit_mc(&b_empty[bs]);
#ifdef CG_TC_INSTRUMENT
                dacc[0] += tm1 - tm0; dacc[1] += CG_CLK() - tm1; dacc[8] += 1;   // issue / commit time of one weight stage
#endif
              }
              __syncwarp();
              accum = 1;
              if (++bs == BS) { bs = 0; bph ^= 1; }
            }
            if (elect_one()) umma2_commit_mc(&s_empty[ss]);
            __syncwarp();
            if (++ss == SS) { ss = 0; sph ^= 1; }
          }
        }
        if (elect_one()) umma2_commit_mc(&tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    const int lq = warp & 3, half = (warp - 2) >> 2;
    uint8_t* stg = stage_buf + (warp - 2) * kStgBytes;
    uint8_t* stg_partner = stage_buf + ((warp - 2) ^ 4) * kStgBytes;
    const uint32_t tempty_leader[2] = {mapa_u32(smem_u32(&tempty[0]), 0), mapa_u32(smem_u32(&tempty[1]), 0)};
    int it = 0, last_nt = -1;
    for (int t = pair; t < total_tiles; t += npairs, ++it) {
      const int mt = t % P.m_tiles;
      const int rest = t / P.m_tiles;
      const int nt = rest % P.n_tiles;
      const int phase = rest / P.n_tiles;
      const int blk = mt * 2 + (int)rank;   // blocks past the end map to samples >= B and are masked
      EpiRows R;
      const int bq = blk / P.blocks_per_sample;
      const int sft = (p.ps_out || EPI == EPI_PS_MASK) ? p.ps_shift[(bq < p.B ? bq : 0) / p.ps_group_b] : 0;
      bool need_x = false;
      if (EPI == EPI_PS_MASK) {   // host guarantees single-sample blocks (Q % 128 == 0)
        const int bi = blk % P.blocks_per_sample;
        epi_rows_ps(p, bq, bi * 128, phase, sft, lq, lane, R);
        need_x = (sft > 0 && bi == P.blocks_per_sample - 1) || (sft < 0 && bi == 0);
      } else if (PT) epi_rows(p, blk * P.bpt, 0, P.rpt_log2, phase, lq, lane, R);
      else epi_rows(p, bq, (blk % P.blocks_per_sample) * 128, 7, phase, lq, lane, R);
      if (nt != last_nt) { epi_load_bias<EPI>(p, bias_s, nt * BN, BN, threadIdx.x - 64); last_nt = nt; }
      const int acc = it & 1;
      long long tq = CG_CLK();
      mbar_wait(&tfull[acc], ((uint32_t)it >> 1) & 1);
      if (warp == 2) CG_DBG_ADD(5, tq);
      tq = CG_CLK();
      tc_fence_after();
      if (!(P.dbg_flags & 1))   // timing experiment bit 1: no epilogue
        epilogue_block<EPI>(p, tmem_base + acc * P.acc_stride, BN, nt * BN, R, bias_s, stg, stg_partner, lq, lane, half,
                            (blk % P.blocks_per_sample) * 128, sft, need_x, bq);
      if (warp == 2) CG_DBG_ADD(6, tq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
    }
  }

  if (CG_DBG_ON && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[7] = CG_CLK() - t_kernel0;
  if (warp == 1 || warp == 2) CG_DBG_FLUSH;
  tc_fence_before();
  cluster_sync_all();          // nobody may exit (or free TMEM) while the peer can still signal / read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

// =============================================================================================
// wgrad_tc: one CTA per (m-tile = 2 units of 64 channels, n-tile, row split)
// =============================================================================================
struct WgTcParams {
  WgParams p;
  int BN;             // multiple of 64, <= 256
  int n_origin;       // first output column of this launch's n-tiles
  int n_tiles, m_tiles, units, mblocks;   // mblocks = Mp/64, units = nseg*mblocks
  int rpt, bpt, chunks_per_sample, total_chunks, chunks_per_split, splits, stages;
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmP,
                const __grid_constant__ WgTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const WgParams& p = P.p;
  const int BN = P.BN;
  const int stages = P.stages;
  const int stage_bytes = kABytes + BN * 128;   // A: 2 blocks x (64 rows x 128 B); B: BN/64 blocks
  uint8_t* tiles = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item
  int w = blockIdx.x;
  const int split = w % P.splits; w /= P.splits;
  const int nt = w % P.n_tiles;
  const int mt = w / P.n_tiles;
  const int ch_begin = split * P.chunks_per_split;
  int ch_end = ch_begin + P.chunks_per_split;
  if (ch_end > P.total_chunks) ch_end = P.total_chunks;
  const int nchunks = ch_end - ch_begin;   // >= 1 by construction
  const int n_begin = P.n_origin + nt * BN;
  int unit[2] = {mt * 2, mt * 2 + 1};
  const bool unit1_ok = unit[1] < P.units;
  if (!unit1_ok) unit[1] = unit[0];

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmS);
    prefetch_tmap(&tmP);
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0;
    uint32_t ph = 0;
    int shift[2], scol[2];
    for (int u = 0; u < 2; ++u) {
      const int seg = unit[u] / P.mblocks, mb = unit[u] % P.mblocks;
      shift[u] = p.shift[seg];
      scol[u] = p.scol[seg] + mb * 64;
    }
    for (int ch = ch_begin; ch < ch_end; ++ch) {
      int b0, q0;
      if (P.chunks_per_sample > 0) { b0 = ch / P.chunks_per_sample; q0 = (ch % P.chunks_per_sample) * 64; }
      else { b0 = ch * P.bpt; q0 = 0; }
      mbar_wait(&empty[stage], ph ^ 1);
      if (elect_one()) {
        uint8_t* sa = tiles + (size_t)stage * stage_bytes;
        mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
        tma_load_3d(sa, &tmS, &full[stage], scol[0], q0 + shift[0], b0);
        tma_load_3d(sa + 8192, &tmS, &full[stage], scol[1], q0 + shift[1], b0);
        for (int j = 0; j < BN / 64; ++j)
          tma_load_3d(sa + kABytes + j * 8192, &tmP, &full[stage], n_begin + j * 64, q0, b0);
      }
      if (++stage == stages) { stage = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, BN, 1, 1);
    // MN-major, SWIZZLE_128B: LBO = stride between 64-channel blocks (8192 B), SBO = stride between 8-row groups
    const uint32_t hi = desc_hi(1024);
    const uint32_t lo0 = desc_lo(smem_u32(tiles), 8192);
    const uint32_t stage_step = (uint32_t)stage_bytes >> 4;
    int stage = 0;
    uint32_t ph = 0;
    for (int ci = 0; ci < nchunks; ++ci) {
      mbar_wait(&full[stage], ph);
      tc_fence_after();
      const uint32_t a_lo = lo0 + stage * stage_step;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)   // K = 16 rows = 2 groups of 8 rows = 2048 bytes per step
          umma_bf16_lohi(tmem_base, a_lo + 128 * k, a_lo + (kABytes >> 4) + 128 * k, hi, idesc, (uint32_t)(ci | k));
        umma_commit(&empty[stage]);
        if (ci == nchunks - 1) umma_commit(&tfull[0]);
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; ph ^= 1; }
    }
  } else {
    const int lq = warp & 3;
    const int r = lq * 32 + lane;       // tile row: unit r/64, channel r%64
    const int u = r >> 6;
    const int seg = unit[u] / P.mblocks, mb = unit[u] % P.mblocks;
    const int m = mb * 64 + (r & 63);
    const bool row_ok = (u == 0 || unit1_ok) && m < p.m_real;
    float* dst = p.dW + ((long long)seg * p.m_real + m) * p.n_real;
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n_begin + c0 + j;
          if (n < p.n_real) atomicAdd(dst + n, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// =============================================================================================
// wgrad2_tc: slab-reuse weight gradient (time rows per sample >= 64).
//   work item = (tap group g, 64-channel block mb, tap subset, n-tile, row split). Per 64-row chunk ONE slab
//   S[b, q0+min_shift .. q0+64+max_shift, acol_g + 64*mb .. +64] is loaded; accumulator a holds taps (ta, tb):
//   its M = 128 operand is the slab read at row offsets shift(ta) and shift(tb) -- the second 64-channel MN block
//   is reached through the descriptor's LBO = (shift(tb) - shift(ta)) rows. All accumulators share the P tile.
// =============================================================================================
struct Wg2Params {
  WgParams p;
  int BN, n_origin, n_tiles;
  int mblocks, ngroups, nsub[2], taps_per_cta;       // nsub[g] = tap subsets of group g
  int items_per_split, splits;                        // items = sum_g mblocks * nsub[g] * n_tiles
  int chunks_per_sample, total_chunks, chunks_per_split;
  int box_rows, slab_bytes, stages;
  int g_acol[2], g_min_shift[2], g_nseg[2];
  unsigned char g_rel[2][32];                         // shift - min_shift of tap s of group g
  unsigned char g_seg[2][32];                         // original tap index (output slice)
  int swap;   // 1: operands exchanged (slab = output gradient with negated shifts, P tile = layer input at column g_acol):
              //    accumulator rows = output channels, columns = input channels, stored transposed
  long long* dbg;   // optional role cycle counters of CTA 0 (instrumented build, CG_TC_TIMING=1)
  // CTA-pair variant (wgrad2p_tc_kernel): two work items of one (row split, n-tile) form a cluster; every item has
  // pair_ntaps taps whose row offsets relative to the item's first tap are rel_pat[], so ONE descriptor set serves both CTAs
  int pb_half, pair_ntaps, p_rows_total;
  unsigned char rel_pat[16];
  int bulk;         // 1: the epilogue stages accumulator rows in the (then idle) stage ring and adds them to dW with one
                    //    bulk reduce per row segment (TMA engine) instead of red.global.v4 per 16 bytes, whose LSU issue rate
                    //    (~1.3 clk per lane-op) made the tail 12-15 % of the kernel
};

__device__ __forceinline__ void bulk_red_add_f32(float* gdst, uint32_t src_s, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(src_s), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Epilogue of the weight-gradient kernels: accumulators (TMEM) -> gradient buffer, added to the other row splits' partial sums.
__device__ __forceinline__ void wgrad2_epilogue(const Wg2Params& P, uint8_t* tiles, uint32_t tmem_base, int BN, int n_begin, int g,
                                                int mb, int tap0, int ntaps, int nacc, int lq, int half, int r, int m, int lane) {
  const WgParams& p = P.p;
  const bool vec_ok = (p.n_real & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.dW) & 15) == 0);
  (void)lane;
  if (P.bulk && P.swap) {
    // transposed store dW[seg][n][m]: stage each accumulator as [tap half][n][64 channels] (lanes = consecutive m, so
    // the 4-byte stores are conflict-free), then one 256-byte bulk reduce-add per (tap, input channel) row
    const uint32_t buf = smem_u32(tiles);
    const int e = (int)threadIdx.x - 64;
    int mcnt = p.m_real - mb * 64;
    if (mcnt > 64) mcnt = 64;
    for (int a = 0; a < nacc; ++a) {
      if (a) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      for (int c0 = half * 32; c0 < BN; c0 += 64) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + a * BN + c0, v);
        tmem_ld_wait();
        const uint32_t dst_s = buf + (uint32_t)(((r >> 6) * BN + c0) * 256 + (r & 63) * 4);
#pragma unroll
        for (int j = 0; j < 32; ++j) asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst_s + j * 256), "r"(v[j]) : "memory");
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int idx = e; idx < 2 * BN; idx += 256) {
        const int th = idx / BN, n = idx - th * BN;
        const int ti = tap0 + 2 * a + th;
        if (ti < tap0 + ntaps && n_begin + n < p.n_real && mcnt > 0) {
          const int seg = P.g_seg[g][ti];
          bulk_red_add_f32(p.dW + ((long long)seg * p.n_real + n_begin + n) * p.m_real + mb * 64,
                           buf + (uint32_t)(th * BN + n) * 256u, (uint32_t)mcnt * 4u);
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else if (P.bulk) {
    // row r of every accumulator: this thread's half of the columns -> padded row buffer (conflict-free 16-byte
    // stores) -> one bulk reduce-add of the whole segment into dW
    const int nch = BN >> 5, cb = half ? (nch + 1) >> 1 : 0, ce = half ? nch : (nch + 1) >> 1;
    const uint32_t rowbuf = smem_u32(tiles) + (uint32_t)r * (uint32_t)(BN * 4 + 16);
    const int n0 = n_begin + cb * 32;
    int n1 = n_begin + ce * 32;
    if (n1 > p.n_real) n1 = p.n_real;
    for (int a = 0; a < nacc; ++a) {
      const int ti = tap0 + 2 * a + (r >> 6);
      const bool row_ok = ti < tap0 + ntaps && m < p.m_real;
      const int seg = P.g_seg[g][ti < 32 ? ti : 0];
      float* dst = p.dW + ((long long)seg * p.m_real + m) * p.n_real;
      if (a) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // my previous segment has left shared memory
      for (int c = cb; c < ce; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + a * BN + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) st_shared_v4(rowbuf + c * 128 + j * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      if (row_ok && n1 > n0) {
        fence_proxy_async();
        bulk_red_add_f32(dst + n0, rowbuf + cb * 128, (uint32_t)(n1 - n0) * 4u);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else
  for (int a = 0; a < nacc; ++a) {
    const int ti = tap0 + 2 * a + (r >> 6);
    const bool row_ok = ti < tap0 + ntaps && m < p.m_real;
    const int seg = P.g_seg[g][ti < 32 ? ti : 0];
    float* dst = p.dW + ((long long)seg * p.m_real + m) * p.n_real;
    for (int c0 = half * 32; c0 < BN; c0 += 64) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + a * BN + c0, v);
      tmem_ld_wait();
      if (row_ok && P.swap) {   // transposed store: element (m = output channel, n = input channel) -> dW[seg][n][m];
                                // consecutive lanes hold consecutive m, so every warp instruction is one 128-byte line
        const int n0 = n_begin + c0;
        float* dcol = p.dW + ((long long)seg * p.n_real + n0) * p.m_real + m;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.n_real) atomicAdd(dcol + (long long)j * p.m_real, __uint_as_float(v[j]));
      } else if (row_ok) {
        const int n0 = n_begin + c0;
        if (vec_ok && n0 + 32 <= p.n_real) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst + n0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                       __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.n_real) atomicAdd(dst + n0 + j, __uint_as_float(v[j]));
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kWg2Threads, 1)
wgrad2_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmP,
                 const __grid_constant__ Wg2Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = P.BN;
  const int stages = P.stages;
  const int p_blocks = (BN + 63) >> 6;                  // 64-column MN-major blocks of the P tile (BN = 160 loads 3, uses 2.5)
  const int stage_bytes = P.slab_bytes + p_blocks * 8192;
  uint8_t* tiles = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- decode the work item
  int w = blockIdx.x;
  const int split = w / P.items_per_split;
  w -= split * P.items_per_split;
  const int nt = w % P.n_tiles; w /= P.n_tiles;
  int g = 0;
  if (w >= P.mblocks * P.nsub[0]) { g = 1; w -= P.mblocks * P.nsub[0]; }
  const int sub = w % P.nsub[g];
  const int mb = w / P.nsub[g];
  const int tap0 = sub * P.taps_per_cta;
  int ntaps = P.g_nseg[g] - tap0;
  if (ntaps > P.taps_per_cta) ntaps = P.taps_per_cta;
  const int nacc = (ntaps + 1) / 2;
  const int ch_begin = split * P.chunks_per_split;
  int ch_end = ch_begin + P.chunks_per_split;
  if (ch_end > P.total_chunks) ch_end = P.total_chunks;
  const int nchunks = ch_end - ch_begin;
  const int n_begin = P.n_origin + nt * BN;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmS);
    prefetch_tmap(&tmP);
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long e_t0_w = 0, e_t1_w = 0;   // epilogue timestamps (instrumented build)
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    int stage = 0;
    uint32_t ph = 0;
    const int scol = (P.swap ? 0 : P.g_acol[g]) + mb * 64;
    const int pcol = (P.swap ? P.g_acol[g] : 0) + n_begin;
    for (int ch = ch_begin; ch < ch_end; ++ch) {
      const int b0 = ch / P.chunks_per_sample, q0 = (ch % P.chunks_per_sample) * 64;
      mbar_wait(&empty[stage], ph ^ 1);
      if (elect_one()) {
        uint8_t* sa = tiles + (size_t)stage * stage_bytes;
        mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
        tma_load_3d(sa, &tmS, &full[stage], scol, q0 + P.g_min_shift[g], b0);
        for (int j = 0; j < p_blocks; ++j)
          tma_load_3d(sa + P.slab_bytes + j * 8192, &tmP, &full[stage], pcol + j * 64, q0, b0);
      }
      if (++stage == stages) { stage = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, BN, 1, 1);
    const uint32_t hi = desc_hi(1024);
    const uint32_t lo0 = (smem_u32(tiles) >> 4) & 0x3FFF;
    const uint32_t stage_step = (uint32_t)stage_bytes >> 4;
    const uint32_t b_off = ((uint32_t)P.slab_bytes >> 4) | (512u << 16);   // P tile: LBO = 8192 B between 64-channel blocks
    // per accumulator: row offset of the first tap (16-byte units) | LBO = distance to the second tap
    uint32_t a_off[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int ta = tap0 + 2 * a, tb = (2 * a + 1 < ntaps) ? ta + 1 : ta;
      const uint32_t ra = a < nacc ? P.g_rel[g][ta] : 0, rb = a < nacc ? P.g_rel[g][tb] : 0;
      a_off[a] = ra * 8 + (((rb - ra) * 8) << 16);
    }
    int stage = 0;
    uint32_t ph = 0;
    long long w_wait = 0, w_issue = 0;
    const long long w_t0 = CG_CLK();
    for (int ci = 0; ci < nchunks; ++ci) {
      const long long tq0 = CG_CLK();
      mbar_wait(&full[stage], ph);
      tc_fence_after();
      const long long tq1 = CG_CLK();
      w_wait += tq1 - tq0;
      const uint32_t s_lo = lo0 + stage * stage_step;
      if (elect_one()) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          if (a < nacc) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // K = 16 rows = 2048 bytes per step
              umma_bf16_lohi(tmem_base + a * BN, s_lo + a_off[a] + 128 * k, s_lo + b_off + 128 * k, hi, idesc,
                             (uint32_t)(ci | k));
          }
        }
        umma_commit(&empty[stage]);
        if (ci == nchunks - 1) umma_commit(&tfull[0]);
      }
      __syncwarp();
      w_issue += CG_CLK() - tq1;
      if (++stage == stages) { stage = 0; ph ^= 1; }
    }
    if (CG_DBG_ON && blockIdx.x == 0 && lane == 0) {
      P.dbg[0] = w_wait; P.dbg[1] = w_issue; P.dbg[2] = CG_CLK() - w_t0; P.dbg[3] = nchunks; P.dbg[4] = nacc;
    }
  } else {
    const int lq = warp & 3, half = (warp - 2) >> 2;
    const int r = lq * 32 + lane;       // accumulator row: tap (r / 64) of the pair, channel r % 64
    e_t0_w = CG_CLK();
    const int m = mb * 64 + (r & 63);
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    e_t1_w = CG_CLK();
    wgrad2_epilogue(P, tiles, tmem_base, BN, n_begin, g, mb, tap0, ntaps, nacc, lq, half, r, m, lane);
  }
  if (CG_DBG_ON && blockIdx.x == 0 && threadIdx.x == 64) { P.dbg[5] = CG_CLK() - e_t1_w; P.dbg[6] = e_t1_w - e_t0_w; }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =============================================================================================
// wgrad2p_tc: CTA-pair (cta_group::2) weight gradient. The single-CTA kernel is shared-memory bound: per M128 x N x K16 MMA
// it reads A 4 KB + B 32*N B while TMA fills the stage ring, (reads + fills) / 128 B/clk per SM fits the measured loop
// times of all four critic layers within 6 % (N = 128: 162 B/clk needed, N = 256: 137). Here two work items that share
// the row split and the n-tile -- hence the P tile (output-gradient chunk) -- form a cluster: each CTA keeps its own slab,
// accumulators and epilogue, but loads and reads only HALF of the P tile (BN/2 columns); the leader issues M = 256 MMAs.
// Barriers as in rsgemm3: full on the leader (both CTAs' TMA loads complete_tx there), empty / accumulator-full
// multicast-committed to both CTAs.
// =============================================================================================
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWg2Threads, 1)
wgrad2p_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmP,
                  const __grid_constant__ Wg2Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = P.BN;
  const int stages = P.stages;
  const int stage_bytes = P.slab_bytes + P.pb_half * 8192;   // own slab + own half of the P tile
  uint8_t* tiles = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  // ---- decode the work item: consecutive items of a split pair up; the pair shares (split, n-tile)
  int w = blockIdx.x;
  const int split = w / P.items_per_split;
  w -= split * P.items_per_split;
  const int pq = w >> 1;
  const int nt = pq % P.n_tiles;
  w = (pq / P.n_tiles) * 2 + (w & 1);
  int g = 0;
  if (w >= P.mblocks * P.nsub[0]) { g = 1; w -= P.mblocks * P.nsub[0]; }
  const int sub = w % P.nsub[g];
  const int mb = w / P.nsub[g];
  const int tap0 = sub * P.taps_per_cta;
  const int ntaps = P.pair_ntaps;
  const int nacc = (ntaps + 1) / 2;
  const int ch_begin = split * P.chunks_per_split;
  int ch_end = ch_begin + P.chunks_per_split;
  if (ch_end > P.total_chunks) ch_end = P.total_chunks;
  const int nchunks = ch_end - ch_begin;
  const int n_begin = P.n_origin + nt * BN;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmS);
    prefetch_tmap(&tmP);
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long e_t0_w = 0, e_t1_w = 0;
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    int stage = 0;
    uint32_t ph = 0;
    const int scol = (P.swap ? 0 : P.g_acol[g]) + mb * 64;
    const int pcol = (P.swap ? P.g_acol[g] : 0) + n_begin + (int)rank * (BN / 2);
    const int row0 = P.g_min_shift[g] + P.g_rel[g][tap0];      // the slab starts at this item's first tap
    for (int ch = ch_begin; ch < ch_end; ++ch) {
      const int b0 = ch / P.chunks_per_sample, q0 = (ch % P.chunks_per_sample) * 64;
      mbar_wait(&empty[stage], ph ^ 1);
      if (elect_one()) {
        uint8_t* sa = tiles + (size_t)stage * stage_bytes;
        if (rank == 0) mbar_expect_tx(&full[stage], (uint32_t)(2 * stage_bytes));
        const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
        tma2_load_3d(sa, &tmS, bar, scol, q0 + row0, b0);
        for (int j = 0; j < P.pb_half; ++j)
          tma2_load_3d(sa + P.slab_bytes + j * 8192, &tmP, bar, pcol + j * 64, q0, b0);
      }
      if (++stage == stages) { stage = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc(256, BN, 1, 1);
      const uint32_t hi = desc_hi(1024);
      const uint32_t lo0 = (smem_u32(tiles) >> 4) & 0x3FFF;
      const uint32_t stage_step = (uint32_t)stage_bytes >> 4;
      const uint32_t b_off = ((uint32_t)P.slab_bytes >> 4) | (512u << 16);   // P half tile: LBO = 8192 B between 64-channel blocks
      uint32_t a_off[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int ia = 2 * a, ib = (2 * a + 1 < ntaps) ? 2 * a + 1 : 2 * a;
        const uint32_t ra = a < nacc ? P.rel_pat[ia] : 0, rb = a < nacc ? P.rel_pat[ib] : 0;
        a_off[a] = ra * 8 + (((rb - ra) * 8) << 16);
      }
      int stage = 0;
      uint32_t ph = 0;
      long long w_wait = 0, w_issue = 0;
      const long long w_t0 = CG_CLK();
      for (int ci = 0; ci < nchunks; ++ci) {
        const long long tq0 = CG_CLK();
        mbar_wait(&full[stage], ph);
        tc_fence_after();
        const long long tq1 = CG_CLK();
        w_wait += tq1 - tq0;
        const uint32_t s_lo = lo0 + stage * stage_step;
        if (elect_one()) {
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            if (a < nacc) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma2_bf16_lohi(tmem_base + a * BN, s_lo + a_off[a] + 128 * k, s_lo + b_off + 128 * k, hi, idesc,
                                (uint32_t)(ci | k));
            }
          }
          umma2_commit_mc(&empty[stage]);
          if (ci == nchunks - 1) umma2_commit_mc(&tfull[0]);
        }
        __syncwarp();
        w_issue += CG_CLK() - tq1;
        if (++stage == stages) { stage = 0; ph ^= 1; }
      }
      if (CG_DBG_ON && blockIdx.x == 0 && lane == 0) {
        P.dbg[0] = w_wait; P.dbg[1] = w_issue; P.dbg[2] = CG_CLK() - w_t0; P.dbg[3] = nchunks; P.dbg[4] = nacc;
      }
    }
  } else {
    const int lq = warp & 3, half = (warp - 2) >> 2;
    const int r = lq * 32 + lane;       // accumulator row: tap (r / 64) of the pair, channel r % 64
    e_t0_w = CG_CLK();
    const int m = mb * 64 + (r & 63);
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    e_t1_w = CG_CLK();
    wgrad2_epilogue(P, tiles, tmem_base, BN, n_begin, g, mb, tap0, ntaps, nacc, lq, half, r, m, lane);
  }
  if (CG_DBG_ON && blockIdx.x == 0 && threadIdx.x == 64) { P.dbg[5] = CG_CLK() - e_t1_w; P.dbg[6] = e_t1_w - e_t0_w; }
  tc_fence_before();
  cluster_sync_all();          // nobody may exit (or free TMEM) while the peer can still signal / read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
  PFN_encodeTiled encode = nullptr;
  std::map<std::tuple<const void*, long long, long long, long long, long long, int, int, int>, CUtensorMap> cache;
  int sm_count = 148;
  int max_smem = 0;
  bool force_v1 = false;   // CG_TC_V1=1: per-tap boxes everywhere (A/B comparison)
  const bool use_pair = true;   // the cta_group::2 kernel serves every layer with >= 128 time rows per sample
  bool pair_short = true;  // CG_TC_PAIR_SHORT=0: per-tap single-CTA kernel for layers with < 128 time rows
  bool ghead_attr_set = false;
  long long *dbg_buf = nullptr, *dbg_buf3 = nullptr;   // role cycle counters (instrumented builds only), per context / device
  std::string err;
};

extern int cg_tc_set_err(const char* msg);   // defined in cg_engine.cu

static inline int tc_init(TcState* s) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return cg_tc_set_err("cuTensorMapEncodeTiled not available from the driver");
  s->encode = (PFN_encodeTiled)fn;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return cg_tc_set_err("cudaGetDeviceProperties failed");
  if (prop.major != 10) return cg_tc_set_err("the bf16 tensor-core path needs an sm_100 device (tcgen05/TMEM)");
  s->sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("CG_SM_LIMIT")) {   // experiment: leave SMs free for a concurrent collective
    const int lim = atoi(e) & ~1;
    if (lim >= 2 && lim < s->sm_count) s->sm_count = lim;
  }
  // 2 KB short of the opt-in maximum: leaves room on every SM for a second, tiny CTA (1 KB of system-reserved shared
  // memory each) -- the data-parallel peer-memory reduction kernel runs beside the persistent GEMM CTAs
  s->max_smem = (int)prop.sharedMemPerBlockOptin - 2048;
  if (const char* e = getenv("CG_TC_V1")) s->force_v1 = atoi(e) != 0;
  if (const char* e = getenv("CG_TC_PAIR_SHORT")) s->pair_short = atoi(e) != 0;
  bool ok = true;
#define CG_SET_SMEM(E)                                                                                                         \
  ok = ok && cudaFuncSetAttribute(tc::rsgemm_tc_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) == cudaSuccess && \
       cudaFuncSetAttribute(tc::rsgemm3_tc_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) == cudaSuccess;
  CG_SET_SMEM(EPI_NONE) CG_SET_SMEM(EPI_BIAS) CG_SET_SMEM(EPI_BIAS_LRELU) CG_SET_SMEM(EPI_MASK) CG_SET_SMEM(EPI_BIAS_SIGMOID)
  ok = ok && cudaFuncSetAttribute(tc::rsgemm3_tc_kernel<EPI_BIAS_LN_LRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) == cudaSuccess &&
       cudaFuncSetAttribute(tc::rsgemm3_tc_kernel<EPI_PS_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) == cudaSuccess;
#undef CG_SET_SMEM
  if (!ok ||
      cudaFuncSetAttribute(tc::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(tc::wgrad2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(tc::wgrad2p_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) != cudaSuccess)
    return cg_tc_set_err("cudaFuncSetAttribute(max dynamic smem) failed");
  return 0;
}
static inline void tc_destroy(TcState* s) {
  s->cache.clear();
  if (s->dbg_buf) cudaFree(s->dbg_buf);
  if (s->dbg_buf3) cudaFree(s->dbg_buf3);
  s->dbg_buf = s->dbg_buf3 = nullptr;
}

// 3-D map over a bf16 tensor viewed as (batch, rows, cols) with box (64 cols, box_rows, box_batch), SWIZZLE_128B
static inline int tc_get_map3(TcState* s, const void* base, long long cols, long long rows, long long batch,
                              long long row_stride, long long batch_stride, int box_rows, int box_batch,
                              CUtensorMap* out) {
  auto key = std::make_tuple(base, cols, rows, batch, row_stride * 1000003LL + batch_stride, box_rows, box_batch, 3);
  auto it = s->cache.find(key);
  if (it != s->cache.end()) { *out = it->second; return 0; }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstr[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_batch};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = s->encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(3d) failed: %d (cols %lld rows %lld batch %lld rs %lld bs %lld box %d x %d)",
             (int)r, cols, rows, batch, row_stride, batch_stride, box_rows, box_batch);
    return cg_tc_set_err(buf);
  }
  s->cache[key] = m;
  *out = m;
  return 0;
}
static inline int tc_get_map2(TcState* s, const void* base, long long cols, long long rows, long long row_stride,
                              int box_rows, CUtensorMap* out) {
  auto key = std::make_tuple(base, cols, rows, 0LL, row_stride, box_rows, 0, 2);
  auto it = s->cache.find(key);
  if (it != s->cache.end()) { *out = it->second; return 0; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)row_stride * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = s->encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(2d) failed: %d (cols %lld rows %lld rs %lld box %d)", (int)r,
             cols, rows, row_stride, box_rows);
    return cg_tc_set_err(buf);
  }
  s->cache[key] = m;
  *out = m;
  return 0;
}

// kernel launch with the programmatic-dependent-launch attribute (see griddep_wait)
template <typename... KArgs, typename... Args>
static inline void tc_launch(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
  static const bool pdl = getenv("CG_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors surface through cudaGetLastError in post_launch
}

static inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// rows of one 128-row M tile must come from whole samples (Q | 128) or one sample (128 | Q)
static inline bool tc_rsgemm_supported(const RsParams& p) {
  if (p.Kc % 64 || p.N % 64) return false;
  if (!((p.Q >= 128 && p.Q % 128 == 0) || (p.Q < 128 && is_pow2(p.Q)))) return false;
  if (p.a_rows != p.Q) return false;
  return true;
}

static inline int tc_pick_bn(int N) {
  for (int bn = 256; bn >= 64; bn -= 32)
    if (N % bn == 0) return bn;
  return 64;
}

// ---- v2 (slab reuse) -------------------------------------------------------------------------------
static inline bool tc_rsgemm2_supported(const RsParams& p) {
  if (p.Q < 128 || p.Q % 128) return false;
  for (int ph = 0; ph < p.seg.nphase; ++ph) {
    int acols[4], na = 0, cnt[4] = {0, 0, 0, 0};
    for (int s = 0; s < p.seg.nseg[ph]; ++s) {
      int g = -1;
      for (int i = 0; i < na; ++i) if (acols[i] == p.seg.acol[ph][s]) g = i;
      if (g < 0) { if (na == 4) return false; acols[na] = p.seg.acol[ph][s]; g = na++; }
      if (++cnt[g] > 32) return false;
    }
  }
  return true;
}

// layer-norm can be fused into the conv epilogue when the slab kernels apply and one n-tile covers the channel row
static inline bool tc_ln_fusable(const RsParams& p) { return p.N <= 256 && p.N % 32 == 0; }

// ---- v3: CTA-pair kernel -----------------------------------------------------------------------------
static inline int tc_rsgemm3_launch(TcState* s, const RsParams& p, cudaStream_t stream) {
  tc::RsTc2Params P;
  memset(&P, 0, sizeof(P));
  P.p = p;
  P.per_tap = p.Q < 128 ? 1 : 0;
  if (p.epi == EPI_PS_MASK && (P.per_tap || (p.seg.nphase != 2 && !p.merged_phases) || !p.mask || p.ps_w != 2 * p.Q))
    return cg_tc_set_err("rsgemm3_tc: EPI_PS_MASK needs single-sample blocks (Q % 128 == 0) and the two-phase data-gradient form");
  int span = 0;
  for (int ph = 0; ph < p.seg.nphase; ++ph) {
    int na = 0;
    for (int sg = 0; sg < p.seg.nseg[ph]; ++sg) {
      int g = -1;
      for (int i = 0; i < na; ++i) if (P.grp[ph][i].acol == p.seg.acol[ph][sg]) g = i;
      if (g < 0) { g = na++; P.grp[ph][g].acol = p.seg.acol[ph][sg]; P.grp[ph][g].min_shift = 1 << 20; P.grp[ph][g].nseg = 0; }
      if (p.seg.shift[ph][sg] < P.grp[ph][g].min_shift) P.grp[ph][g].min_shift = p.seg.shift[ph][sg];
    }
    P.ngroups[ph] = na;
    for (int sg = 0; sg < p.seg.nseg[ph]; ++sg) {
      int g = 0;
      for (int i = 0; i < na; ++i) if (P.grp[ph][i].acol == p.seg.acol[ph][sg]) g = i;
      tc::SlabGroup& G = P.grp[ph][g];
      const int rel = p.seg.shift[ph][sg] - G.min_shift;
      G.shift_rel[G.nseg] = (unsigned char)rel;
      G.wk[G.nseg] = p.seg.wk[ph][sg];
      G.nseg++;
      if (rel > span) span = rel;
    }
  }
  P.box_rows = (128 + span + 7) / 8 * 8;
  if (P.box_rows > 256) return cg_tc_set_err("rsgemm3_tc: tap span too large for one TMA box");
  P.box_bytes = P.box_rows * 128;
  P.blocks_per_sample = p.Q / 128;
  P.total_blocks = p.B * P.blocks_per_sample;
  if (P.per_tap) {   // whole samples per 128-row block
    P.rpt = p.Q; P.bpt = 128 / p.Q;
    P.rpt_log2 = 0;
    while ((1 << P.rpt_log2) < P.rpt) ++P.rpt_log2;
    P.blocks_per_sample = 1;
    P.total_blocks = (p.B + P.bpt - 1) / P.bpt;
    P.box_bytes = tc::kABytes;
  }
  P.kchunks = p.Kc / 64;
  P.MB = 1;
  P.m_tiles = (P.total_blocks + 1) / 2;
  const int npairs_max = s->sm_count / 2;
  // widest BN (<= 256, multiple of 32) that divides N, unless a narrower one gives fewer waves over the CTA pairs
  double best = 1e30;
  int bestBN = 64;
  for (int BN = 256; BN >= 64; BN -= 32) {
    if (p.N % BN) continue;
    const long long tiles = (long long)p.seg.nphase * (p.N / BN) * P.m_tiles;
    const double waves = (double)((tiles + npairs_max - 1) / npairs_max);
    const double math = BN * 2.0;                                            // 4 MMAs x BN/2 clk per SM
    const double smem_clk = (4.0 * (4096 + BN * 16) + BN * 64 + P.box_bytes / 12.0) / 128.0;
    const double per = math > smem_clk ? math : smem_clk;
    const double cost = waves * per;
    if (cost < best) { best = cost; bestBN = BN; }
  }
  if (const char* e = getenv("CG_TC_BN")) { const int bn = atoi(e); if (bn >= 32 && bn % 32 == 0 && p.N % bn == 0) bestBN = bn; }   // tuning experiments
  if (p.epi == EPI_BIAS_LN_LRELU) bestBN = p.N;   // the epilogue needs whole channel rows
  P.BN = bestBN;
  P.n_tiles = p.N / P.BN;
  // One n-tile whose real channels end well before the 64-channel padding (102 of 128): issue N = 112 MMAs (any multiple
  // of 16 is a legal cta_group::2 shape) instead of multiplying 26 all-zero weight rows -- 12% fewer tensor cycles on the
  // generator's last conv-transpose and the critic's first data gradient.
  if (P.n_tiles == 1 && !p.row_pairs && !p.merged_phases && !getenv("CG_TC_NO_TRIM")) {
    const int trimmed = (p.n_real + 15) / 16 * 16;
    if (trimmed >= 32 && trimmed < P.BN) P.BN = trimmed;
  }
  P.acc_stride = (P.BN + 31) / 32 * 32;
  P.double_acc = 1;
  P.slab_bytes = P.box_bytes;
  int groups_per_tile = 0;
  for (int ph = 0; ph < p.seg.nphase; ++ph) if (P.ngroups[ph] > groups_per_tile) groups_per_tile = P.ngroups[ph];
  groups_per_tile *= P.kchunks;
  P.slab_stages = groups_per_tile <= 2 ? 4 : (groups_per_tile <= 4 ? 3 : 2);   // short K loops: prefetch the next tile's slabs
  if (p.row_pairs) P.slab_stages = 3;   // four short tap groups per 64-channel chunk
  if (const char* e = getenv("CG_TC_SST")) { const int v = atoi(e); if (v >= 2 && v <= 6 && !P.per_tap) P.slab_stages = v; }   // tuning experiments
  // taps per weight stage: enough MMA time per stage (4 MMAs x BN/2 clk per tap) to cover a cross-CTA barrier round-trip
  int stage_clk = 1000;
  if (const char* e = getenv("CG_TC_STAGE_CLK")) stage_clk = atoi(e);
  P.tps = (stage_clk + 2 * P.BN - 1) / (2 * P.BN);
  if (P.tps < 1) P.tps = 1;
  if (P.tps > 6) P.tps = P.BN <= 64 ? 12 : 6;   // N = 64: a whole 12-tap parity group per weight stage (measured 140 -> 131 us on conv1)
  if (P.per_tap && P.tps > 3) P.tps = 3;
  if (p.row_pairs || p.merged_phases) {   // few, equal weight stages per tap group (6-7 taps each): measured 108 -> 97 us on
    int most = 1;                         // conv1 forward (tps 4 -> 7)
    for (int ph = 0; ph < p.seg.nphase; ++ph)
      for (int g = 0; g < P.ngroups[ph]; ++g) if (P.grp[ph][g].nseg > most) most = P.grp[ph][g].nseg;
    const int nst = (most + 7) / 8;
    P.tps = (most + nst - 1) / nst;
  }
  if (const char* e = getenv("CG_TC_TPS")) P.tps = atoi(e);
  if (const char* e = getenv("CG_TC_NK")) P.dbg_nk = atoi(e);
  if (const char* e = getenv("CG_TC_DBG")) P.dbg_flags = atoi(e);
  if (P.per_tap) { P.slab_bytes = P.tps * tc::kABytes; P.slab_stages = 3; }
  const int b_bytes = P.tps * P.BN * 64;
  int bst = (s->max_smem - 1024 - 512 - tc::kEpiSmem - P.slab_stages * P.slab_bytes) / b_bytes;
  if (bst > 10) bst = 10;
  if (bst < 2) return cg_tc_set_err("rsgemm3_tc: not enough shared memory");
  P.b_stages = bst;
  P.dbg = nullptr;
  if (CG_TC_INSTRUMENTED && getenv("CG_TC_TIMING") != nullptr) {
    if (!s->dbg_buf3) cudaMalloc(&s->dbg_buf3, 16 * sizeof(long long));
    cudaMemsetAsync(s->dbg_buf3, 0, 16 * sizeof(long long), stream);
    P.dbg = s->dbg_buf3;
  }
  CUtensorMap tmA, tmW;
  if (P.per_tap) { if (tc_get_map3(s, p.A, p.a_rs, p.a_rows, p.B, p.a_rs, p.a_bs, P.rpt, P.bpt, &tmA)) return 1; }
  else if (tc_get_map3(s, p.A, p.a_rs, p.a_rows, p.B, p.a_rs, p.a_bs, P.box_rows, 1, &tmA)) return 1;
  if (tc_get_map2(s, p.W, p.w_ld, p.N, p.w_ld, P.BN / 2, &tmW)) return 1;
  const int total = p.seg.nphase * P.n_tiles * P.m_tiles;
  const int npairs = total < npairs_max ? total : npairs_max;
  const int grid = 2 * npairs;
  const size_t smem = (size_t)P.slab_stages * P.slab_bytes + (size_t)bst * b_bytes + 1024 + 512 + tc::kEpiSmem;
  switch (p.epi) {
    case EPI_NONE: tc_launch(tc::rsgemm3_tc_kernel<EPI_NONE>, grid, tc::kThreads, smem, stream, tmA, tmW, P); break;
    case EPI_BIAS: tc_launch(tc::rsgemm3_tc_kernel<EPI_BIAS>, grid, tc::kThreads, smem, stream, tmA, tmW, P); break;
    case EPI_BIAS_LRELU: tc_launch(tc::rsgemm3_tc_kernel<EPI_BIAS_LRELU>, grid, tc::kThreads, smem, stream, tmA, tmW, P); break;
    case EPI_MASK: tc_launch(tc::rsgemm3_tc_kernel<EPI_MASK>, grid, tc::kThreads, smem, stream, tmA, tmW, P); break;
    case EPI_BIAS_LN_LRELU: tc_launch(tc::rsgemm3_tc_kernel<EPI_BIAS_LN_LRELU>, grid, tc::kThreads, smem, stream, tmA, tmW, P); break;
    case EPI_PS_MASK: tc_launch(tc::rsgemm3_tc_kernel<EPI_PS_MASK>, grid, tc::kThreads, smem, stream, tmA, tmW, P); break;
    default: tc_launch(tc::rsgemm3_tc_kernel<EPI_BIAS_SIGMOID>, grid, tc::kThreads, smem, stream, tmA, tmW, P); break;
  }
  if (P.dbg) {
    long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, s->dbg_buf3, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc3 timing] B=%d Q=%d N=%d Kc=%d epi=%d | BN=%d tps=%d tiles=%d pairs=%d sst=%d bst=%d pt=%d | total %lld clk | mma wait tempty %lld/%lld s_full %lld/%lld b_full %lld/%lld issue %lld commit %lld /%lld | epi wait %lld/%lld work %lld/%lld\n",
            p.B, p.Q, p.N, p.Kc, p.epi, P.BN, P.tps, total, npairs, P.slab_stages, bst, P.per_tap, h[7], h[2], h[10], h[3], h[11], h[4], h[12], h[0], h[1], h[8], h[5], h[13], h[6], h[14]);
  }
  return 0;
}

static inline int tc_rsgemm_launch(TcState* s, const RsParams& p, cudaStream_t stream) {
  if (!s->force_v1 && s->use_pair && (tc_rsgemm2_supported(p) || (s->pair_short && p.Q < 128 && p.B * p.Q >= 256 &&
                                                                (double)p.B * p.Q * p.N * p.Kc * (p.seg.nseg[0] + p.seg.nseg[1]) >= 8e9)))
    return tc_rsgemm3_launch(s, p, stream);
  if (p.epi == EPI_BIAS_LN_LRELU || p.epi == EPI_PS_MASK || p.ps_out)
    return cg_tc_set_err("rsgemm_tc: fused layer-norm / PhaseShuffle epilogues need the CTA-pair kernel");
  tc::RsTcParams P;
  P.p = p;
  P.BN = tc_pick_bn(p.N);
  P.n_tiles = p.N / P.BN;
  if (p.Q >= 128) { P.rpt = 128; P.bpt = 1; P.tiles_per_sample = p.Q / 128; P.m_tiles = p.B * P.tiles_per_sample; }
  else { P.rpt = p.Q; P.bpt = 128 / p.Q; P.tiles_per_sample = 0; P.m_tiles = (p.B + P.bpt - 1) / P.bpt; }
  P.kchunks = p.Kc / 64;
  P.x_shift = 0; P.x_baseoff = 0;
  if (const char* e = getenv("CG_TC_DBG")) P.x_baseoff = atoi(e);   // timing experiments: 1 no stores, 2 no W loads, 4 no A loads
  const int a_rows_box = P.rpt;
  P.a_bytes = tc::kABytes;
  const int stage_bytes = P.a_bytes + P.BN * 128;
  int stages = (s->max_smem - 1024 - 512 - tc::kEpiSmem) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return cg_tc_set_err("rsgemm_tc: not enough shared memory for 2 stages");
  P.stages = stages;
  CUtensorMap tmA, tmW;
  // A viewed as (B, a_rows, a_rs): column extent = a_rs (the strided view covers both parities)
  if (tc_get_map3(s, p.A, p.a_rs, p.a_rows, p.B, p.a_rs, p.a_bs, a_rows_box, P.bpt, &tmA)) return 1;
  if (tc_get_map2(s, p.W, p.w_ld, p.N, p.w_ld, P.BN, &tmW)) return 1;
  const int total = p.seg.nphase * P.n_tiles * P.m_tiles;
  const int grid = total < s->sm_count ? total : s->sm_count;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 512 + tc::kEpiSmem;
  P.dbg = nullptr;
  if (CG_TC_INSTRUMENTED && getenv("CG_TC_TIMING") != nullptr) {
    if (!s->dbg_buf) cudaMalloc(&s->dbg_buf, 16 * sizeof(long long));
    cudaMemsetAsync(s->dbg_buf, 0, 16 * sizeof(long long), stream);
    P.dbg = s->dbg_buf;
  }
  const long long t_host0 = 0;
  (void)t_host0;
  switch (p.epi) {
    case EPI_NONE: tc::rsgemm_tc_kernel<EPI_NONE><<<grid, tc::kThreads, smem, stream>>>(tmA, tmW, P); break;
    case EPI_BIAS: tc::rsgemm_tc_kernel<EPI_BIAS><<<grid, tc::kThreads, smem, stream>>>(tmA, tmW, P); break;
    case EPI_BIAS_LRELU: tc::rsgemm_tc_kernel<EPI_BIAS_LRELU><<<grid, tc::kThreads, smem, stream>>>(tmA, tmW, P); break;
    case EPI_MASK: tc::rsgemm_tc_kernel<EPI_MASK><<<grid, tc::kThreads, smem, stream>>>(tmA, tmW, P); break;
    default: tc::rsgemm_tc_kernel<EPI_BIAS_SIGMOID><<<grid, tc::kThreads, smem, stream>>>(tmA, tmW, P); break;
  }
  if (P.dbg) {
    long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, s->dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc timing] N=%d BN=%d Q=%d tiles=%d grid=%d stages=%d | producer wait-empty %lld clk/%lld | mma wait-tempty %lld/%lld, wait-full %lld/%lld | epi wait-tfull %lld, work %lld over %lld tiles\n",
            p.N, P.BN, p.Q, total, grid, stages, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
  }
  return 0;
}

static inline bool tc_wgrad_supported(const WgParams& p) {
  if (p.Mp % 64 || p.Np % 64) return false;
  if (!((p.Q >= 64 && p.Q % 64 == 0) || (p.Q < 64 && is_pow2(p.Q)))) return false;
  if (p.s_rows != p.Q) return false;
  return true;
}

// ---- wgrad v2 (slab reuse) -----------------------------------------------------------------------------
static inline bool tc_wgrad2_supported(const WgParams& p) {
  if (p.Q < 64 || p.Q % 64 || p.s_rows != p.Q) return false;
  int cols[2], nc = 0, cnt[2] = {0, 0};
  for (int s = 0; s < p.nseg; ++s) {
    int g = -1;
    for (int i = 0; i < nc; ++i) if (cols[i] == p.scol[s]) g = i;
    if (g < 0) { if (nc == 2) return false; cols[nc] = p.scol[s]; g = nc++; }
    if (++cnt[g] > 32) return false;
  }
  return true;
}

// Output-channel count 64 makes the N = 64 MMA shared-memory bound (A 4 KB + B 2 KB per 128 x 64 x 16 MACs: 48 clk for
// 32 clk of math). Exchanging the operands -- dW[k][ci][co] = sum_q' X[q'][ci] * dY[q' - shift_k][co] -- puts the 64
// output channels of TWO taps on M = 128 and the (>= 128) input channels on N: 8 KB per 128 x 128 x 16 MACs = math rate.
static inline bool tc_wgrad2_swap_pays(const WgParams& p) { return p.Np == 64 && p.Mp >= 128 && p.Mp <= 256 && !getenv("CG_WG_NO_SWAP"); }
static inline WgParams tc_wgrad2_swapped(const WgParams& p) {
  WgParams w = p;
  w.S = p.P; w.s_bs = p.p_bs; w.s_rs = p.p_rs; w.s_rows = p.Q;
  w.P = p.S; w.p_bs = p.s_bs; w.p_rs = p.s_rs;
  w.m_real = p.n_real; w.n_real = p.m_real; w.Mp = p.Np; w.Np = p.Mp;
  for (int k = 0; k < p.nseg; ++k) w.shift[k] = (short)(-p.shift[k]);   // scol[] keeps naming the parity group = P-tile column origin
  return w;
}

// CTA-pair launch of one pass of tc_wgrad2_launch (P holds the pass' tiling; items = work items per row split). Returns
// false when the work items cannot be paired: odd count, tap subsets of different size or spacing, or (operand-swapped
// mode, where the P tile depends on the tap group) a pair that would straddle the two groups.
static inline bool tc_wgrad2_pair_launch(TcState* s, const tc::Wg2Params& P0, const WgParams& p, bool swap, int items,
                                         cudaStream_t stream) {
  if (getenv("CG_WG_NO_PAIR")) return false;
  tc::Wg2Params P = P0;
  const int combos = items / P.n_tiles;
  if (combos % 2 || P.BN % 32 || P.BN > 256) return false;
  if (swap && (P.mblocks * P.nsub[0]) % 2) return false;
  int ntaps0 = -1, span = 0;
  for (int g = 0; g < P.ngroups; ++g)
    for (int sub = 0; sub < P.nsub[g]; ++sub) {
      const int tap0 = sub * P.taps_per_cta;
      int nt = P.g_nseg[g] - tap0;
      if (nt > P.taps_per_cta) nt = P.taps_per_cta;
      if (nt > 16) return false;
      if (ntaps0 < 0) {
        ntaps0 = nt;
        for (int j = 0; j < nt; ++j) {
          P.rel_pat[j] = (unsigned char)(P.g_rel[g][tap0 + j] - P.g_rel[g][tap0]);
          if (P.rel_pat[j] > span) span = P.rel_pat[j];
        }
      } else {
        if (nt != ntaps0) return false;
        for (int j = 0; j < nt; ++j)
          if (P.rel_pat[j] != (unsigned char)(P.g_rel[g][tap0 + j] - P.g_rel[g][tap0])) return false;
      }
    }
  if (ntaps0 <= 0) return false;
  P.pair_ntaps = ntaps0;
  P.pb_half = (P.BN / 2 + 63) / 64;
  P.box_rows = (64 + span + 7) / 8 * 8;
  P.slab_bytes = P.box_rows * 128;
  CUtensorMap tmS, tmP;
  if (tc_get_map3(s, p.S, p.s_rs, p.s_rows, p.B, p.s_rs, p.s_bs, P.box_rows, 1, &tmS)) return false;
  if (tc_get_map3(s, p.P, p.p_rs, P.p_rows_total, p.B, p.p_rs, p.p_bs, 64, 1, &tmP)) return false;
  int splits = s->sm_count / items;
  if (splits > P.total_chunks) splits = P.total_chunks;
  if (splits < 1) return false;
  P.items_per_split = items;
  P.chunks_per_split = (P.total_chunks + splits - 1) / splits;
  P.splits = (P.total_chunks + P.chunks_per_split - 1) / P.chunks_per_split;
  const int stage_bytes = P.slab_bytes + P.pb_half * 8192;
  int stages = (s->max_smem - 2048) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return false;
  P.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  const bool dw_aligned = (reinterpret_cast<uintptr_t>(p.dW) & 15) == 0 && !getenv("CG_WG_NO_BULK");
  if (swap)
    P.bulk = (dw_aligned && P.BN % 64 == 0 && p.m_real % 4 == 0 && (size_t)2 * P.BN * 256 <= (size_t)stages * stage_bytes) ? 1 : 0;
  else
    P.bulk = (dw_aligned && p.n_real % 4 == 0 && P.n_origin % 4 == 0 &&
              (size_t)128 * (P.BN * 4 + 16) <= (size_t)stages * stage_bytes) ? 1 : 0;
  P.dbg = nullptr;
  if (CG_TC_INSTRUMENTED && getenv("CG_TC_TIMING") != nullptr) {
    if (!s->dbg_buf) cudaMalloc(&s->dbg_buf, 16 * sizeof(long long));
    cudaMemsetAsync(s->dbg_buf, 0, 16 * sizeof(long long), stream);
    P.dbg = s->dbg_buf;
  }
  tc_launch(tc::wgrad2p_tc_kernel, items * P.splits, tc::kWg2Threads, smem, stream, tmS, tmP, P);
  if (P.dbg) {
    long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, s->dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[wg2 timing] PAIR B=%d Q=%d M=%d N=%d swap=%d bulk=%d | BN=%d items=%d splits=%d stages=%d box_rows=%d | chunks %lld nacc %lld | mma: wait full %lld issue+commit %lld loop total %lld | epilogue: wait %lld work %lld\n",
            p.B, p.Q, p.Mp, p.Np, P.swap, P.bulk, P.BN, items, P.splits, stages, P.box_rows, h[3], h[4], h[0], h[1], h[2], h[6], h[5]);
  }
  return true;
}

static inline int tc_wgrad2_launch(TcState* s, const WgParams& p_in, cudaStream_t stream) {
  tc::Wg2Params P;
  memset(&P, 0, sizeof(P));
  const bool swap = tc_wgrad2_swap_pays(p_in);
  const WgParams p = swap ? tc_wgrad2_swapped(p_in) : p_in;
  P.swap = swap ? 1 : 0;
  P.p = p;
  P.mblocks = p.Mp / 64;
  // tap groups (same column offset = same parity), taps sorted by shift so LBO >= 0
  for (int sg = 0; sg < p.nseg; ++sg) {
    int g = -1;
    for (int i = 0; i < P.ngroups; ++i) if (P.g_acol[i] == p.scol[sg]) g = i;
    if (g < 0) { g = P.ngroups++; P.g_acol[g] = p.scol[sg]; P.g_min_shift[g] = 1 << 20; }
    if (p.shift[sg] < P.g_min_shift[g]) P.g_min_shift[g] = p.shift[sg];
  }
  int span = 0;
  for (int g = 0; g < P.ngroups; ++g) {
    int idx[32], n = 0;
    for (int sg = 0; sg < p.nseg; ++sg) if (p.scol[sg] == P.g_acol[g]) idx[n++] = sg;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j)
        if (p.shift[idx[j]] < p.shift[idx[i]]) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
    P.g_nseg[g] = n;
    for (int i = 0; i < n; ++i) {
      P.g_rel[g][i] = (unsigned char)(p.shift[idx[i]] - P.g_min_shift[g]);
      P.g_seg[g][i] = (unsigned char)idx[i];
      if (P.g_rel[g][i] > span) span = P.g_rel[g][i];
    }
  }
  P.box_rows = (64 + span + 7) / 8 * 8;
  if (P.box_rows > 256) return cg_tc_set_err("wgrad2_tc: tap span too large for one TMA box");
  P.slab_bytes = P.box_rows * 128;
  P.chunks_per_sample = p.Q / 64;
  P.total_chunks = p.B * P.chunks_per_sample;
  CUtensorMap tmS, tmP;
  if (tc_get_map3(s, p.S, p.s_rs, p.s_rows, p.B, p.s_rs, p.s_bs, P.box_rows, 1, &tmS)) return 1;
  if (tc_get_map3(s, p.P, p.p_rs, swap ? p_in.s_rows : p.Q, p.B, p.p_rs, p.p_bs, 64, 1, &tmP)) return 1;
  P.p_rows_total = swap ? p_in.s_rows : p.Q;
  // 256 < Np <= 512 with Np/2 a multiple of 32 (320 -> 2 x 160): two equal n-tiles in ONE launch instead of a 256-wide
  // launch plus a 64-wide remainder launch that runs shared-memory bound (measured 33% tensor-pipe activity)
  const bool halves = p.Np > 256 && p.Np <= 512 && (p.Np / 2) % 32 == 0 && !getenv("CG_WG_NO_HALVES");
  for (int pass = 0; pass < 2; ++pass) {
    const int full_tiles = p.Np / 256, rem = p.Np % 256;
    if (halves) { if (pass) continue; P.BN = p.Np / 2; P.n_origin = 0; P.n_tiles = 2; }
    else if (pass == 0) { if (!full_tiles) continue; P.BN = 256; P.n_origin = 0; P.n_tiles = full_tiles; }
    else { if (!rem) continue; P.BN = rem; P.n_origin = full_tiles * 256; P.n_tiles = 1; }
    int max_acc = 512 / P.BN;
    if (max_acc > 8) max_acc = 8;
    P.taps_per_cta = 2 * max_acc;
    {   // same number of tap subsets, but balanced (12 taps, 8 per CTA: 8 + 4 -> 6 + 6): every CTA loads the same bytes
      int most = 0;
      for (int g = 0; g < P.ngroups; ++g) if (P.g_nseg[g] > most) most = P.g_nseg[g];
      const int nsub = (most + P.taps_per_cta - 1) / P.taps_per_cta;
      const int even = ((most + nsub - 1) / nsub + 1) / 2 * 2;
      if (even < P.taps_per_cta && !getenv("CG_WG_NO_BALANCE")) P.taps_per_cta = even;
    }
    int items = 0;
    for (int g = 0; g < 2; ++g) {
      P.nsub[g] = g < P.ngroups ? (P.g_nseg[g] + P.taps_per_cta - 1) / P.taps_per_cta : 0;
      items += P.mblocks * P.nsub[g] * P.n_tiles;
    }
    if (P.nsub[0] == 0) return cg_tc_set_err("wgrad2_tc: empty tap group");
    P.items_per_split = items;
    // whole waves only: items x splits <= waves x SMs (a ceil here used to leave a third, nearly empty wave)
    int waves = 1;
    if (const char* e = getenv("CG_WG_WAVES")) waves = atoi(e);
    int splits = (waves * s->sm_count) / items;
    if (splits > P.total_chunks) splits = P.total_chunks;
    if (splits < 1) splits = 1;
    P.chunks_per_split = (P.total_chunks + splits - 1) / splits;
    P.splits = (P.total_chunks + P.chunks_per_split - 1) / P.chunks_per_split;
    if (tc_wgrad2_pair_launch(s, P, p, swap, items, stream)) continue;
    const int stage_bytes = P.slab_bytes + ((P.BN + 63) / 64) * 8192;
    int stages = (s->max_smem - 2048) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return cg_tc_set_err("wgrad2_tc: not enough shared memory");
    P.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    // bulk-reduce epilogue: direct (row-major) stores, 16-byte aligned row segments, the row staging area (128 padded
    // rows) fits in the stage ring
    const bool dw_aligned = (reinterpret_cast<uintptr_t>(p.dW) & 15) == 0 && !getenv("CG_WG_NO_BULK");
    if (swap)   // transposed store: 64-channel rows of dW[seg][n][:]
      P.bulk = (dw_aligned && P.BN % 64 == 0 && p.m_real % 4 == 0 && (size_t)2 * P.BN * 256 <= (size_t)stages * stage_bytes &&
                !getenv("CG_WG_NO_BULK_SWAP")) ? 1 : 0;
    else
      P.bulk = (dw_aligned && P.BN % 32 == 0 && p.n_real % 4 == 0 && P.n_origin % 4 == 0 &&
                (size_t)128 * (P.BN * 4 + 16) <= (size_t)stages * stage_bytes) ? 1 : 0;
    P.dbg = nullptr;
    if (CG_TC_INSTRUMENTED && getenv("CG_TC_TIMING") != nullptr) {
      if (!s->dbg_buf) cudaMalloc(&s->dbg_buf, 16 * sizeof(long long));
      cudaMemsetAsync(s->dbg_buf, 0, 16 * sizeof(long long), stream);
      P.dbg = s->dbg_buf;
    }
    tc_launch(tc::wgrad2_tc_kernel, items * P.splits, tc::kWg2Threads, smem, stream, tmS, tmP, P);
    if (P.dbg) {
      long long h[16];
      cudaStreamSynchronize(stream);
      cudaMemcpy(h, s->dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[wg2 timing] B=%d Q=%d M=%d N=%d swap=%d bulk=%d | BN=%d items=%d splits=%d stages=%d | chunks %lld nacc %lld | mma: wait full %lld issue+commit %lld loop total %lld | epilogue: wait %lld work %lld\n",
              p.B, p.Q, p.Mp, p.Np, P.swap, P.bulk, P.BN, items, P.splits, stages, h[3], h[4], h[0], h[1], h[2], h[6], h[5]);
    }
  }
  return 0;
}

static inline int tc_wgrad_launch(TcState* s, const WgParams& p, cudaStream_t stream) {
  if (!s->force_v1 && tc_wgrad2_supported(p)) return tc_wgrad2_launch(s, p, stream);
  tc::WgTcParams P;
  P.p = p;
  P.mblocks = p.Mp / 64;
  P.units = p.nseg * P.mblocks;
  P.m_tiles = (P.units + 1) / 2;
  P.n_tiles = (p.Np + 255) / 256;
  if (p.Q >= 64) { P.rpt = 64; P.bpt = 1; P.chunks_per_sample = p.Q / 64; P.total_chunks = p.B * P.chunks_per_sample; }
  else { P.rpt = p.Q; P.bpt = 64 / p.Q; P.chunks_per_sample = 0; P.total_chunks = (p.B + P.bpt - 1) / P.bpt; }
  CUtensorMap tmS, tmP;
  if (tc_get_map3(s, p.S, p.s_rs, p.s_rows, p.B, p.s_rs, p.s_bs, P.rpt, P.bpt, &tmS)) return 1;
  if (tc_get_map3(s, p.P, p.p_rs, p.Q, p.B, p.p_rs, p.p_bs, P.rpt, P.bpt, &tmP)) return 1;
  // n-tiles are 256 wide, plus one narrower remainder tile: one launch per distinct BN
  for (int pass = 0; pass < 2; ++pass) {
    const int full_tiles = p.Np / 256, rem = p.Np % 256;
    if (pass == 0) { if (!full_tiles) continue; P.BN = 256; P.n_origin = 0; P.n_tiles = full_tiles; }
    else { if (!rem) continue; P.BN = rem; P.n_origin = full_tiles * 256; P.n_tiles = 1; }
    const int tiles = P.m_tiles * P.n_tiles;
    int splits = (2 * s->sm_count + tiles - 1) / tiles;
    if (splits > P.total_chunks) splits = P.total_chunks;
    if (splits < 1) splits = 1;
    P.chunks_per_split = (P.total_chunks + splits - 1) / splits;
    P.splits = (P.total_chunks + P.chunks_per_split - 1) / P.chunks_per_split;
    const int stage_bytes = tc::kABytes + P.BN * 128;
    int stages = (s->max_smem - 2048) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return cg_tc_set_err("wgrad_tc: not enough shared memory");
    P.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    tc::wgrad_tc_kernel<<<tiles * P.splits, tc::kWgThreads, smem, stream>>>(tmS, tmP, P);
  }
  return 0;
}
