// Generator output head for the bf16 path (sm_100a): per-timestep Dense(C) + sigmoid (calciumgan.py:96-101), fused
// with the WGAN-GP interpolation x_hat = alpha * real + (1 - alpha) * fake (wgan_gp.py:38-41).
//
// The layer is HBM-bound (K = N = 128 per output row: ~100 FLOP/B), so the design goal is bytes, not MMAs:
//   in : activations (rows, 128) bf16 by TMA, optionally the real batch (rows, C) fp32 by one 1-D bulk copy per tile
//   out: any of  fake bf16 (rows, 128)  -> critic input slot "fake"     (TMA tensor store, 128B-swizzled staging)
//                x_hat bf16 (rows, 128) -> critic input slot "x_hat"    (TMA tensor store)
//                fake fp32 (rows, C)    -> FAKE32                       (ONE 1-D bulk store per tile: a 128-row tile of
//                                                                         the unpadded fp32 tensor is contiguous)
// Persistent CTAs, one 128-row tile at a time: warp 0 TMA producer (weights once, activation tiles), warp 1 issues
// the 8 tcgen05.mma (M 128 x N 128 x K 16) of a tile into a double-buffered TMEM accumulator, warps 2-9 are the
// epilogue (two warps per TMEM lane quarter, 64 columns each; thread = output row). The epilogue never touches
// global memory with ld/st: rows are staged in shared memory in their final layout and moved by the async proxy.
#pragma once
#include "cg_kernels_tc.cuh"

namespace tc {

struct GHeadParams {
  const float* bias;    // [C]
  const float* real;    // (rows, C) fp32 or null
  const float* alpha;   // [B] (with real)
  float* out32;         // (rows, C) fp32 or null
  int tiles, L, C, sigmoid, a_stages, has_f, has_x;
};

constexpr int kHeadThreads = 320;
constexpr int kHeadCp = 128;                       // padded channels on both sides of the layer
constexpr int kHeadTileBytes = 128 * kHeadCp * 2;  // one bf16 tile (two 64-channel swizzled boxes)

__device__ __forceinline__ void bulk_load_1d(uint32_t dst_s, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t src_s, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_s, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src_s), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void head_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps
__device__ __forceinline__ void sts_v2f(uint32_t sa, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sa), "f"(a), "f"(b));
}
__device__ __forceinline__ float2 lds_v2f(uint32_t sa) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sa));
  return v;
}

__global__ void __launch_bounds__(kHeadThreads, 1)
ghead_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ GHeadParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int AS = P.a_stages;
  const uint32_t bytes32 = 128u * (uint32_t)P.C * 4u;   // fp32 tile, multiple of 1024 (C even)
  uint8_t* w_s = smem;                                  // [2 chunks][128 n][128 B]
  uint8_t* a_s = w_s + kHeadTileBytes;                  // AS x [2 chunks][128 rows][128 B]
  uint8_t* stgF = a_s + (size_t)AS * kHeadTileBytes;
  uint8_t* stgX = stgF + (P.has_f ? kHeadTileBytes : 0);
  uint8_t* buf32 = stgX + (P.has_x ? kHeadTileBytes : 0);
  float* bias_s = reinterpret_cast<float*>(buf32 + ((P.real || P.out32) ? bytes32 : 0));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(bias_s + kHeadCp);
  uint64_t* a_empty = a_full + 4;
  uint64_t* tfull = a_empty + 4;
  uint64_t* tempty = tfull + 2;
  uint64_t* w_full = tempty + 2;
  uint64_t* r_full = w_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    if (P.has_f) prefetch_tmap(&tmF);
    if (P.has_x) prefetch_tmap(&tmX);
    for (int i = 0; i < AS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    mbar_init(w_full, 1);
    mbar_init(r_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kHeadCp);
  griddep_wait();      // programmatic dependent launch: the prologue above overlaps the predecessor's tail
  griddep_launch();
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kHeadCp) {
    const int n = threadIdx.x - 64;
    bias_s[n] = n < P.C ? __ldg(&P.bias[n]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_full, (uint32_t)kHeadTileBytes);
      for (int kc = 0; kc < 2; ++kc) tma_load_2d(w_s + kc * (kHeadTileBytes / 2), &tmW, w_full, kc * 64, 0);
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < P.tiles; t += gridDim.x) {
      mbar_wait(&a_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&a_full[s], (uint32_t)kHeadTileBytes);
        for (int kc = 0; kc < 2; ++kc)
          tma_load_2d(a_s + (size_t)s * kHeadTileBytes + kc * (kHeadTileBytes / 2), &tmA, &a_full[s], kc * 64, t * 128);
      }
      __syncwarp();
      if (++s == AS) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, kHeadCp, 0, 0);
    const uint32_t hi = desc_hi(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(a_s), 16), w_lo0 = desc_lo(smem_u32(w_s), 16);
    mbar_wait(w_full, 0);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < P.tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty[acc], (((uint32_t)it >> 1) & 1) ^ 1);
      mbar_wait(&a_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * kHeadCp;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(d_tmem, a_lo0 + (uint32_t)(s * kHeadTileBytes + kc * (kHeadTileBytes / 2)) / 16 + 2 * k,
                           w_lo0 + (uint32_t)(kc * (kHeadTileBytes / 2)) / 16 + 2 * k, hi, idesc, (uint32_t)(kc | k));
        umma_commit(&a_empty[s]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++s == AS) { s = 0; ph ^= 1; }
    }
  } else {
    const int lq = warp & 3;              // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;     // which 64 output columns
    const int c0 = half * 64;
    const int row = lq * 32 + lane;
    const bool E0 = threadIdx.x == 64;
    const uint32_t buf32_s = smem_u32(buf32);
    const uint32_t rowF = smem_u32(stgF) + half * (kHeadTileBytes / 2) + row * 128;
    const uint32_t rowX = smem_u32(stgX) + half * (kHeadTileBytes / 2) + row * 128;
    const uint32_t row32 = buf32_s + (uint32_t)(row * P.C + c0) * 4;
    const int sw = lane & 7;
    if (P.real && E0 && blockIdx.x < P.tiles) {
      mbar_expect_tx(r_full, bytes32);
      bulk_load_1d(buf32_s, P.real + (size_t)blockIdx.x * 128 * P.C, bytes32, r_full);
    }
    int it = 0;
    for (int t = blockIdx.x; t < P.tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int tn = t + gridDim.x;
      uint32_t v[64];
      mbar_wait(&tfull[acc], ((uint32_t)it >> 1) & 1);
      tc_fence_after();
      {
        uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        const uint32_t ta = tmem_base + acc * kHeadCp + ((uint32_t)(lq * 32) << 16) + c0;
        tmem_ld32(ta, v0);
        tmem_ld32(ta + 32, v1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
#pragma unroll
      for (int j4 = 0; j4 < 16; ++j4) {
        const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + j4 * 4);
        const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = __uint_as_float(v[j4 * 4 + e]) + bv[e];
          if (P.sigmoid) x = sigmoid_fast(x);
          v[j4 * 4 + e] = __float_as_uint(c0 + j4 * 4 + e < P.C ? x : 0.f);   // padded channels are exact zeros
        }
      }
      uint32_t px[32];
      if (P.has_x) {
        const float a = __ldg(&P.alpha[(int)(((long long)t * 128) / P.L)]);
        mbar_wait(r_full, (uint32_t)it & 1);
#pragma unroll
        for (int j2 = 0; j2 < 32; ++j2) {
          float2 r = make_float2(0.f, 0.f);
          const bool ok = c0 + 2 * j2 < P.C;
          if (ok) r = lds_v2f(row32 + j2 * 8);
          const float x0 = a * r.x + (1.f - a) * __uint_as_float(v[2 * j2]);
          const float x1 = a * r.y + (1.f - a) * __uint_as_float(v[2 * j2 + 1]);
          px[j2] = ok ? pack_bf16x2(x0, x1) : 0u;
        }
      }
      if (E0) bulk_wait_read0();   // the previous tile's stores have drained the staging buffers
      head_bar();
      if (P.real && !P.out32 && E0 && tn < P.tiles) {   // real tile consumed by everybody: fetch the next one
        mbar_expect_tx(r_full, bytes32);
        bulk_load_1d(buf32_s, P.real + (size_t)tn * 128 * P.C, bytes32, r_full);
      }
      if (P.has_f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
          sts_v4(rowF + ((j ^ sw) << 4), o);
        }
      }
      if (P.has_x) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sts_v4(rowX + ((j ^ sw) << 4), make_uint4(px[j * 4], px[j * 4 + 1], px[j * 4 + 2], px[j * 4 + 3]));
      }
      if (P.out32) {   // in place over the consumed real tile: every element is read and written by the same thread
#pragma unroll
        for (int j2 = 0; j2 < 32; ++j2)
          if (c0 + 2 * j2 < P.C) sts_v2f(row32 + j2 * 8, __uint_as_float(v[2 * j2]), __uint_as_float(v[2 * j2 + 1]));
      }
      fence_proxy_async();
      head_bar();
      if (E0) {
        if (P.has_f)
          for (int kc = 0; kc < 2; ++kc) tma_store_2d(&tmF, smem_u32(stgF) + kc * (kHeadTileBytes / 2), kc * 64, t * 128);
        if (P.has_x)
          for (int kc = 0; kc < 2; ++kc) tma_store_2d(&tmX, smem_u32(stgX) + kc * (kHeadTileBytes / 2), kc * 64, t * 128);
        if (P.out32) bulk_store_1d(P.out32 + (size_t)t * 128 * P.C, buf32_s, bytes32);
        bulk_commit();
        if (P.real && P.out32 && tn < P.tiles) {   // validation path: the buffer is both input and output, so serialise
          bulk_wait_read0();
          mbar_expect_tx(r_full, bytes32);
          bulk_load_1d(buf32_s, P.real + (size_t)tn * 128 * P.C, bytes32, r_full);
        }
      }
    }
    if (E0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kHeadCp);
  }
}

}  // namespace tc

struct GHeadArgs {
  const void* A;        // (rows, 128) bf16 activations
  const void* W;        // (128, 128) bf16 [n][k]
  const float* bias;
  void* fake16;         // (rows, 128) bf16 or null
  void* xhat16;         // (rows, 128) bf16 or null (needs real + alpha)
  float* out32;         // (rows, C) fp32 or null
  const float* real;    // (rows, C) fp32 or null
  const float* alpha;   // [B]
  int B, L, C, Cp, sigmoid;
};

static inline bool tc_ghead_supported(const GHeadArgs& a) {
  if (a.Cp != tc::kHeadCp || a.L % 128 || a.C % 2 || a.C > a.Cp || a.C < 2) return false;
  if (a.xhat16 && (!a.real || !a.alpha)) return false;
  if ((reinterpret_cast<uintptr_t>(a.real) | reinterpret_cast<uintptr_t>(a.out32)) & 15) return false;   // 1-D bulk copies
  return true;
}

static inline int tc_ghead_launch(TcState* s, const GHeadArgs& a, cudaStream_t stream) {
  if (!s->ghead_attr_set) {   // function attributes are per device: once per context, not once per process
    if (cudaFuncSetAttribute(tc::ghead_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_smem) != cudaSuccess)
      return cg_tc_set_err("cudaFuncSetAttribute(ghead_tc_kernel) failed");
    s->ghead_attr_set = true;
  }
  tc::GHeadParams P;
  memset(&P, 0, sizeof(P));
  P.bias = a.bias; P.real = a.xhat16 ? a.real : nullptr; P.alpha = a.alpha; P.out32 = a.out32;
  const long long rows = (long long)a.B * a.L;
  P.tiles = (int)(rows / 128); P.L = a.L; P.C = a.C; P.sigmoid = a.sigmoid;
  P.has_f = a.fake16 != nullptr; P.has_x = a.xhat16 != nullptr;
  const size_t fixed = (size_t)tc::kHeadTileBytes * (1 + P.has_f + P.has_x) + ((P.real || P.out32) ? 512u * a.C : 0u) +
                       tc::kHeadCp * 4 + 256 + 1024;
  int as = (int)(((size_t)s->max_smem - fixed) / tc::kHeadTileBytes);
  if (as > 4) as = 4;
  if (as < 2) return cg_tc_set_err("ghead_tc: not enough shared memory");
  P.a_stages = as;
  CUtensorMap tmA, tmW, tmF, tmX;
  if (tc_get_map2(s, a.A, a.Cp, rows, a.Cp, 128, &tmA)) return 1;
  if (tc_get_map2(s, a.W, a.Cp, a.Cp, a.Cp, 128, &tmW)) return 1;
  tmF = tmA; tmX = tmA;
  if (a.fake16 && tc_get_map2(s, a.fake16, a.Cp, rows, a.Cp, 128, &tmF)) return 1;
  if (a.xhat16 && tc_get_map2(s, a.xhat16, a.Cp, rows, a.Cp, 128, &tmX)) return 1;
  const int grid = P.tiles < s->sm_count ? P.tiles : s->sm_count;
  const size_t smem = fixed + (size_t)as * tc::kHeadTileBytes;
  tc_launch(tc::ghead_tc_kernel, grid, tc::kHeadThreads, smem, stream, tmA, tmW, tmF, tmX, P);
  return 0;
}
