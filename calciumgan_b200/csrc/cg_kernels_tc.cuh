// tcgen05 / TMEM / TMA implicit-GEMM kernels for the bf16 mixed-precision path (sm_100a).
//
//  rsgemm_tc : out[b,q,n] = epi( sum_{tap,c} A[b, q+shift(tap), acol(tap)+c] * W[n, wk(tap)+c] )
//              strided Conv1D fwd, Conv1DTranspose fwd, their data gradients, GP linearised fwd, Dense.
//              A tiles are tap-shifted 3-D TMA boxes (batch is its own dim -> per-sample zero fill),
//              W tiles 2-D TMA boxes, both K-major SWIZZLE_128B; D (128 x BN fp32) lives in TMEM,
//              double buffered; warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
//  wgrad_tc  : dW[tap][m][n] += sum_{b,q} S[b, q+shift(tap), scol(tap)+m] * P[b,q,n]
//              both operands MN-major (reduction runs over time rows), split over rows, fp32 red.add.
#pragma once
#include <cuda.h>

#include <cstdlib>
#include <map>
#include <string>
#include <tuple>

#include "cg_common.cuh"

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

constexpr int kThreads = 192;
constexpr int kABytes = 128 * 128;          // 128 rows x 64 bf16
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;             // columns between the two accumulator buffers

struct RsTcParams {
  RsParams p;
  int BN, n_tiles, m_tiles, rpt, bpt, tiles_per_sample, kchunks, stages;
  int a_bytes, x_shift, x_baseoff;   // experiment: A box loaded x_shift rows early, descriptor advanced by x_shift rows
};

// =============================================================================================
// rsgemm_tc
// =============================================================================================
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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.seg.nphase * P.n_tiles * P.m_tiles;

  if (warp == 0) {
    if (lane == 0) {
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
            mbar_wait(&empty[stage], ph ^ 1);
            uint8_t* sa = tiles + (size_t)stage * stage_bytes;
            mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
            tma_load_3d(sa, &tmA, &full[stage], acol + kc * 64, row - P.x_shift, b0);
            tma_load_2d(sa + P.a_bytes, &tmW, &full[stage], wk + kc * 64, nt * BN);
            if (++stage == stages) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, BN, 0, 0);
      int stage = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int phase = (t / P.m_tiles) / P.n_tiles;
        const int acc = it & 1;
        mbar_wait(&tempty[acc], (((uint32_t)it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        const int kiters = p.seg.nseg[phase] * P.kchunks;
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(&full[stage], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * stage_bytes);
          const uint32_t sa_shift = sa + P.x_shift * 128;
          uint64_t adesc = make_desc(sa_shift, 16, 1024);
          if (P.x_baseoff) adesc |= (uint64_t)((sa_shift >> 7) & 7) << 49;
          const uint64_t bdesc = make_desc(sa + P.a_bytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzle row
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (ki | k) != 0);
          umma_commit(&empty[stage]);
          if (ki == kiters - 1) umma_commit(&tfull[acc]);
          if (++stage == stages) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31  == tile rows
    const int lq = warp & 3;
    const int r = lq * 32 + lane;
    bf16* out = reinterpret_cast<bf16*>(p.out);
    const bf16* mask = reinterpret_cast<const bf16*>(p.mask);
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int mt = t % P.m_tiles;
      const int rest = t / P.m_tiles;
      const int nt = rest % P.n_tiles;
      const int phase = rest / P.n_tiles;
      int b, q;
      if (P.tiles_per_sample > 0) { b = mt / P.tiles_per_sample; q = (mt % P.tiles_per_sample) * 128 + r; }
      else { b = mt * P.bpt + r / P.rpt; q = r % P.rpt; }
      const bool row_ok = b < p.B;
      const int acc = it & 1;
      mbar_wait(&tfull[acc], ((uint32_t)it >> 1) & 1);
      tc_fence_after();
      const long long obase = (long long)b * p.o_bs + (long long)q * p.o_rs + phase * p.o_phase_col;
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + acc * kAccStride + c0, v);
        tmem_ld_wait();
        if (row_ok) {
          const int n0 = nt * BN + c0;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            float x = __uint_as_float(v[j]);
            const bool real = n < p.n_real;
            if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_LRELU || p.epi == EPI_BIAS_SIGMOID) x += real ? __ldg(&p.bias[n]) : 0.f;
            if (p.epi == EPI_BIAS_LRELU) x = lrelu(x);
            if (p.epi == EPI_BIAS_SIGMOID) x = 1.f / (1.f + __expf(-x));
            f[j] = real ? x : 0.f;
          }
          if (p.epi == EPI_MASK) {
            const uint4* mp = reinterpret_cast<const uint4*>(mask + obase + n0);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 mv = mp[j4];
              const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
              for (int w2 = 0; w2 < 4; ++w2) {
                const __nv_bfloat162 hv = *reinterpret_cast<const __nv_bfloat162*>(&mw[w2]);
                f[j4 * 8 + w2 * 2] *= lrelu_slope(__low2float(hv));
                f[j4 * 8 + w2 * 2 + 1] *= lrelu_slope(__high2float(hv));
              }
            }
          }
          if (out) {
            uint4* op = reinterpret_cast<uint4*>(out + obase + n0);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint4 o;
              o.x = pack_bf16x2(f[j4 * 8 + 0], f[j4 * 8 + 1]);
              o.y = pack_bf16x2(f[j4 * 8 + 2], f[j4 * 8 + 3]);
              o.z = pack_bf16x2(f[j4 * 8 + 4], f[j4 * 8 + 5]);
              o.w = pack_bf16x2(f[j4 * 8 + 6], f[j4 * 8 + 7]);
              op[j4] = o;
            }
          }
          if (p.out32) {
            float* o32 = p.out32 + (long long)b * p.o32_bs + (long long)q * p.o32_rs;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.n_real) o32[n0 + j] = f[j];
          }
        }
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
// wgrad_tc: one CTA per (m-tile = 2 units of 64 channels, n-tile, row split)
// =============================================================================================
struct WgTcParams {
  WgParams p;
  int BN;             // multiple of 64, <= 256
  int n_origin;       // first output column of this launch's n-tiles
  int n_tiles, m_tiles, units, mblocks;   // mblocks = Mp/64, units = nseg*mblocks
  int rpt, bpt, chunks_per_sample, total_chunks, chunks_per_split, splits, stages;
};

__global__ void __launch_bounds__(kThreads, 1)
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
    if (lane == 0) {
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
        uint8_t* sa = tiles + (size_t)stage * stage_bytes;
        mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
        tma_load_3d(sa, &tmS, &full[stage], scol[0], q0 + shift[0], b0);
        tma_load_3d(sa + 8192, &tmS, &full[stage], scol[1], q0 + shift[1], b0);
        for (int j = 0; j < BN / 64; ++j)
          tma_load_3d(sa + kABytes + j * 8192, &tmP, &full[stage], n_begin + j * 64, q0, b0);
        if (++stage == stages) { stage = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, BN, 1, 1);
      int stage = 0;
      uint32_t ph = 0;
      for (int ci = 0; ci < nchunks; ++ci) {
        mbar_wait(&full[stage], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + (size_t)stage * stage_bytes);
        // MN-major, SWIZZLE_128B: LBO = stride between 64-channel blocks (8192 B), SBO = stride between 8-row groups
        const uint64_t adesc = make_desc(sa, 8192, 1024);
        const uint64_t bdesc = make_desc(sa + kABytes, 8192, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // K = 16 rows = 2 groups of 8 rows = 2048 bytes per step
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (ci | k) != 0);
        umma_commit(&empty[stage]);
        if (ci == nchunks - 1) umma_commit(&tfull[0]);
        if (++stage == stages) { stage = 0; ph ^= 1; }
      }
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
  s->max_smem = (int)prop.sharedMemPerBlockOptin;
  if (cudaFuncSetAttribute(tc::rsgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(tc::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) != cudaSuccess)
    return cg_tc_set_err("cudaFuncSetAttribute(max dynamic smem) failed");
  return 0;
}
static inline void tc_destroy(TcState* s) { s->cache.clear(); }

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

static inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// rows of one 128-row M tile must come from whole samples (Q | 128) or one sample (128 | Q)
static inline bool tc_rsgemm_supported(const RsParams& p) {
  if (p.Kc % 64 || p.N % 64) return false;
  if (!((p.Q >= 128 && p.Q % 128 == 0) || (p.Q < 128 && is_pow2(p.Q)))) return false;
  if (p.a_rows != p.Q) return false;
  return true;
}

static inline int tc_pick_bn(int N) {
  for (int bn = 256; bn >= 64; bn -= 16)
    if (N % bn == 0) return bn;
  return 64;
}

static inline int tc_rsgemm_launch(TcState* s, const RsParams& p, cudaStream_t stream) {
  tc::RsTcParams P;
  P.p = p;
  P.BN = tc_pick_bn(p.N);
  P.n_tiles = p.N / P.BN;
  if (p.Q >= 128) { P.rpt = 128; P.bpt = 1; P.tiles_per_sample = p.Q / 128; P.m_tiles = p.B * P.tiles_per_sample; }
  else { P.rpt = p.Q; P.bpt = 128 / p.Q; P.tiles_per_sample = 0; P.m_tiles = (p.B + P.bpt - 1) / P.bpt; }
  P.kchunks = p.Kc / 64;
  P.x_shift = 0; P.x_baseoff = 0;
  if (const char* e = getenv("CG_TC_XSHIFT")) P.x_shift = atoi(e);
  if (const char* e = getenv("CG_TC_XBASEOFF")) P.x_baseoff = atoi(e);
  if (P.rpt != 128) P.x_shift = 0;
  const int a_rows_box = P.x_shift ? 128 + ((P.x_shift + 7) / 8) * 8 : P.rpt;
  P.a_bytes = P.x_shift ? a_rows_box * 128 : tc::kABytes;
  const int stage_bytes = P.a_bytes + P.BN * 128;
  int stages = (s->max_smem - 2048) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return cg_tc_set_err("rsgemm_tc: not enough shared memory for 2 stages");
  P.stages = stages;
  CUtensorMap tmA, tmW;
  // A viewed as (B, a_rows, a_rs): column extent = a_rs (the strided view covers both parities)
  if (tc_get_map3(s, p.A, p.a_rs, p.a_rows, p.B, p.a_rs, p.a_bs, a_rows_box, P.bpt, &tmA)) return 1;
  if (tc_get_map2(s, p.W, p.w_ld, p.N, p.w_ld, P.BN, &tmW)) return 1;
  const int total = p.seg.nphase * P.n_tiles * P.m_tiles;
  const int grid = total < s->sm_count ? total : s->sm_count;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  tc::rsgemm_tc_kernel<<<grid, tc::kThreads, smem, stream>>>(tmA, tmW, P);
  return 0;
}

static inline bool tc_wgrad_supported(const WgParams& p) {
  if (p.Mp % 64 || p.Np % 64) return false;
  if (!((p.Q >= 64 && p.Q % 64 == 0) || (p.Q < 64 && is_pow2(p.Q)))) return false;
  if (p.s_rows != p.Q) return false;
  return true;
}

static inline int tc_wgrad_launch(TcState* s, const WgParams& p, cudaStream_t stream) {
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
    tc::wgrad_tc_kernel<<<tiles * P.splits, tc::kThreads, smem, stream>>>(tmS, tmP, P);
  }
  return 0;
}
