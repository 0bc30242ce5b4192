// Shared definitions for the calciumgan_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CG_LRELU_ALPHA 0.3f   // tf.keras.layers.LeakyReLU() default (reference gan/models/utils.py:7)
#define CG_LN_EPS 1e-3f       // tf.keras.layers.LayerNormalization() default (gan/models/calciumgan.py:45)
#define CG_CPAD 64            // channel padding granule (one 128-byte swizzled TMA row of bf16)
#define CG_MAX_SEG 64         // max taps (kernel_size)

typedef __nv_bfloat16 bf16;

// ---- element conversion ------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float to_f(float x) { return x; }
  static __device__ __forceinline__ float from_f(float x) { return x; }
};
template <> struct Elem<bf16> {
  static __device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
  static __device__ __forceinline__ bf16 from_f(float x) { return __float2bfloat16_rn(x); }
};

// load 4 consecutive elements as floats (pointer must be 4-element aligned); nullptr -> zeros
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  if (p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = v[1] = v[2] = v[3] = 0.f;
  }
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
  if (p) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  } else {
    v[0] = v[1] = v[2] = v[3] = 0.f;
  }
}

// Programmatic dependent launch for the memory-bound kernels between the GEMMs: a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may be scheduled while its predecessor drains; pdl_enter() lets ITS
// successor do the same and then blocks until the predecessor has completed and its writes are visible. Must be the first
// statement of every kernel launched that way; a no-op under a plain launch.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : CG_LRELU_ALPHA * x; }
__device__ __forceinline__ float lrelu_slope(float h) { return h > 0.f ? 1.f : CG_LRELU_ALPHA; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// PhaseShuffle index map (reference gan/models/calciumgan.py:117-138, closed form SURVEY §8a):
// out[:, t, :] = x[:, ps_index(t, shift, w), :].  int32, bit-exact.
__host__ __device__ __forceinline__ int ps_index(int t, int shift, int w) {
  int j = t + shift;
  j = j < 0 ? -j : j;
  if (j > w - 1) j = 2 * (w - 1) - j;
  return j;
}

// Scatter form of the same map, as used by the fused conv epilogue (cg_kernels_tc.cuh): source row q is the value of
// output row t1 (direct: t1 + shift == q) and of at most one reflected output row t2; -1 = none.
// Checked against ps_index for every (w, shift) on the host (cg_phase_shuffle_scatter_index, tests/test_boundary_cpu.py).
__host__ __device__ __forceinline__ void ps_scatter_targets(int q, int shift, int w, int& t1, int& t2) {
  t1 = q - shift;
  if (t1 < 0 || t1 >= w) t1 = -1;
  t2 = -1;
  if (shift > 0) { t2 = 2 * (w - 1) - q - shift; if (!(t2 >= 0 && t2 < w && t2 + shift > w - 1)) t2 = -1; }
  else if (shift < 0) { t2 = -q - shift; if (!(t2 >= 0 && t2 < w && t2 + shift < 0)) t2 = -1; }
}
// Adjoint (gradient) of the gather as the fused data-gradient epilogue runs it: accumulator row t goes to output row
// dest = t + shift when that lies in [0, w); a row pushed over an edge (dest = -1) is reflected onto the output row
// of a partner row of the same parity: it deposits its value in exchange slot x_src, the partner adds slot x_par.
// Output rows nobody maps to (x_zero, tested on the row with the same index) are written as zeros.
__host__ __device__ __forceinline__ void ps_adjoint_row(int t, int shift, int w, int& dest, int& x_src, int& x_par,
                                                        bool& x_zero) {
  dest = t + shift;
  if (dest < 0 || dest > w - 1) dest = -1;
  x_src = -1; x_par = -1;
  if (shift > 0) {
    if (t + shift > w - 1) x_src = (w - 1 - t) >> 1;
    const int tp = 2 * (w - 1) - t - 2 * shift;
    if (tp <= w - 1 && tp + shift > w - 1) x_par = (w - 1 - tp) >> 1;
  } else if (shift < 0) {
    if (t + shift < 0) x_src = t >> 1;
    const int tp = -t - 2 * shift;
    if (tp >= 0 && tp + shift < 0) x_par = tp >> 1;
  }
  x_zero = t - shift < 0 || t - shift > w - 1;
}

// ---- implicit-GEMM operand addressing ("row-shift GEMM") -----------------------------------
// out[b, q, phase*o_phase_col + n] = epi( sum_{seg in phase} sum_{c<Kc}
//        A[b, q + shift(seg), acol(seg) + c] * W[n, wk(seg) + c] )      (rows outside [0,a_rows) read 0)
struct SegTable {
  int nphase;
  int nseg[2];
  short shift[2][CG_MAX_SEG];
  int acol[2][CG_MAX_SEG];
  int wk[2][CG_MAX_SEG];
};

// EPI_BIAS_LN_LRELU (tensor-core path only, one n-tile = the whole channel row): out = lrelu(LN(acc + bias)),
// optional aux = acc + bias and per-row mean / rstd for the backward pass
// EPI_PS_MASK (CTA-pair tensor-core kernel only): adjoint of the PhaseShuffle gather fused with the LeakyReLU slope of
// the layer below: out[b, j, :] = slope(mask[b, j, :]) * sum_{t : ps_index(t) = j} acc[b, t, :]   (ps_w / ps_shift fields)
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_LRELU = 2, EPI_MASK = 3, EPI_BIAS_SIGMOID = 4, EPI_BIAS_LN_LRELU = 5, EPI_PS_MASK = 6 };

struct RsParams {
  const void* A; long long a_bs; int a_rs; int a_rows;
  const void* W; int w_ld;
  void* out; long long o_bs; int o_rs; int o_phase_col;
  float* out32; long long o32_bs; int o32_rs;      // optional unpadded fp32 copy (n < n_real)
  const float* bias;                               // fp32, n_real entries
  const void* mask;                                // EPI_MASK: same indexing as out
  const float* gamma; const float* beta;           // EPI_BIAS_LN_LRELU: fp32, n_real entries
  float* mu; float* rstd;                          //   optional per-row statistics (row = (b*Q + q)*nphase + phase)
  void* aux;                                       //   optional pre-norm copy, same indexing as out
  // fused PhaseShuffle (slab kernels, strided-conv form): ps_out[b, t, :] = result[b, ps_index(t, shift[b / ps_group_b]), :]
  void* ps_out; int ps_w; int ps_group_b; int ps_shift[4];
  // Row-pair form of a strided convolution with 64 output channels (tensor-core path): GEMM row i holds the outputs of
  // time steps 2i and 2i+1 side by side (N = 128 = [channels of 2i | channels of 2i+1]), the input is read through the
  // (B, L/4, 4*Cp) view and the weights are two tap-shifted copies of the kernel. An N = 64 MMA reads 4 KB of
  // activations for 32 clk of math (shared-memory bound); N = 128 does twice the math on the same 4 KB.
  int row_pairs;                                   // 1: PhaseShuffle scatter / bias index fold the column back to a channel
  // Merged output phases of a transposed-form GEMM with 64 output channels (tensor-core path): both phases read the same
  // input rows, so one N = 128 MMA per input window computes [phase 0 | phase 1] with the two phases' taps stacked in the
  // weight tile (13 windows instead of 2 x 12 taps). Column half = output phase; o_phase_col keeps the phase offset.
  int merged_phases;
  float flop_scale;                                // algorithmic / issued FLOPs (the shifted copies add zero taps); 0 = 1
  int dbg;                                         // timing experiments only
  double* sumsq;                                   // optional: sumsq[b] += sum of squares of the fp32 results of sample b
  int B, Q, N, n_real, Kc, k_real, epi;   // k_real: unpadded channels per tap (algorithmic FLOPs only)
  SegTable seg;
};

// dW[seg][m][n] += sum_{b,q} S[b, q + shift(seg), scol(seg) + m] * P[b, q, n]
struct WgParams {
  const void* S; long long s_bs; int s_rs; int s_rows;
  const void* P; long long p_bs; int p_rs;
  float* dW; int m_real, n_real;
  int B, Q, Mp, Np;
  int nseg;
  int rows_per_split;
  short shift[CG_MAX_SEG];
  int scol[CG_MAX_SEG];
};
