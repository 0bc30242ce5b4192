// CUDA-core kernels of the calciumgan_b200 engine: the fp32 path (reference non-mixed-precision
// mode, parity rel <= 1e-4) and every memory-bound glue op of both precisions.
// All tensors are channels-last (batch, time, channels_padded); T = float | bf16.
#pragma once
#include "cg_common.cuh"

// ---- 16-byte vector helpers: V = 16/sizeof(T) elements (4 fp32 or 8 bf16) ---------------------------------
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };
template <typename T> __device__ __forceinline__ void vload(const T* p, float (&v)[Vec16<T>::N]);
template <> __device__ __forceinline__ void vload<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void vload<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    v[2 * i] = __low2float(h); v[2 * i + 1] = __high2float(h);
  }
}
template <typename T> __device__ __forceinline__ void vstore(T* p, const float (&v)[Vec16<T>::N]);
template <> __device__ __forceinline__ void vstore<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void vstore<bf16>(bf16* p, const float (&v)[8]) {
  uint4 t;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  t.x = w[0]; t.y = w[1]; t.z = w[2]; t.w = w[3];
  *reinterpret_cast<uint4*>(p) = t;
}

// =============================================================================================
// Row-shift implicit GEMM on CUDA cores (fp32 accumulate). 64x64 tile, 256 threads, 4x4/thread.
// Used for: strided Conv1D fwd (reference calciumgan.py:145-185), Conv1DTranspose fwd
// (models/utils.py:79-89), their data gradients, the GP linearised forward, per-timestep Dense.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) rsgemm_simt_kernel(const __grid_constant__ RsParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int phase = blockIdx.z;
  const long long M = (long long)p.B * p.Q;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  const int lrow = tid >> 2;
  const int lc = (tid & 3) * 4;
  const long long r = m0 + lrow;
  const bool rvalid = r < M;
  const int lb = rvalid ? (int)(r / p.Q) : 0;
  const int lq = rvalid ? (int)(r % p.Q) : 0;
  const T* Ab = reinterpret_cast<const T*>(p.A) + (long long)lb * p.a_bs;
  const T* Wr = reinterpret_cast<const T*>(p.W) + (long long)(n0 + lrow) * p.w_ld;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ty = tid >> 4, tx = tid & 15;
  const int nseg = p.seg.nseg[phase];

  for (int s = 0; s < nseg; ++s) {
    const int srow = lq + p.seg.shift[phase][s];
    const bool av = rvalid && srow >= 0 && srow < p.a_rows;
    const T* ap = Ab + (long long)srow * p.a_rs + p.seg.acol[phase][s];
    const T* wp = Wr + p.seg.wk[phase][s];
    for (int c0 = 0; c0 < p.Kc; c0 += BK) {
      float a4[4], w4[4];
      load4<T>(av ? ap + c0 + lc : nullptr, a4);
      load4<T>(wp + c0 + lc, w4);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        As[lc + i][lrow] = a4[i];
        Bs[lc + i][lrow] = w4[i];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 av4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[4] = {av4.x, av4.y, av4.z, av4.w};
        const float b[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  T* out = reinterpret_cast<T*>(p.out);
  const T* mask = reinterpret_cast<const T*>(p.mask);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long rr = m0 + ty * 4 + i;
    if (rr >= M) continue;
    const int b = (int)(rr / p.Q), q = (int)(rr % p.Q);
    const long long obase = (long long)b * p.o_bs + (long long)q * p.o_rs + phase * p.o_phase_col;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      float v = acc[i][j];
      const bool real = n < p.n_real;
      if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_LRELU || p.epi == EPI_BIAS_SIGMOID)
        v += real ? p.bias[n] : 0.f;
      if (p.epi == EPI_BIAS_LRELU) v = lrelu(v);
      if (p.epi == EPI_BIAS_SIGMOID) v = 1.f / (1.f + __expf(-v));
      if (p.epi == EPI_MASK) v *= lrelu_slope(Elem<T>::to_f(mask[obase + n]));
      if (!real) v = 0.f;
      if (out) out[obase + n] = Elem<T>::from_f(v);
      if (p.out32 && real) p.out32[(long long)b * p.o32_bs + (long long)q * p.o32_rs + n] = v;
    }
  }
}

// =============================================================================================
// Weight gradient on CUDA cores: dW[seg][m][n] += sum_rows S[row + shift, scol + m] * P[row, n].
// grid: x = mtiles*ntiles, y = seg, z = row split; fp32 atomics into the flat gradient buffer.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const __grid_constant__ WgParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int ntiles = p.Np / BN;
  const int m0 = (blockIdx.x / ntiles) * BM;
  const int n0 = (blockIdx.x % ntiles) * BN;
  const int seg = blockIdx.y;
  const long long R = (long long)p.B * p.Q;
  const long long r_begin = (long long)blockIdx.z * p.rows_per_split;
  long long r_end = r_begin + p.rows_per_split;
  if (r_end > R) r_end = R;
  const int shift = p.shift[seg];
  const int scol = p.scol[seg];

  const int lr = tid >> 4;
  const int lc = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const T* S = reinterpret_cast<const T*>(p.S);
  const T* P = reinterpret_cast<const T*>(p.P);
  for (long long r0 = r_begin; r0 < r_end; r0 += BK) {
    const long long r = r0 + lr;
    float a4[4], b4[4];
    const T* sp = nullptr;
    const T* pp = nullptr;
    if (r < r_end) {
      const int b = (int)(r / p.Q), q = (int)(r % p.Q);
      const int srow = q + shift;
      if (srow >= 0 && srow < p.s_rows) sp = S + (long long)b * p.s_bs + (long long)srow * p.s_rs + scol + m0 + lc;
      pp = P + (long long)b * p.p_bs + (long long)q * p.p_rs + n0 + lc;
    }
    load4<T>(sp, a4);
    load4<T>(pp, b4);
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lr][lc]) = make_float4(a4[0], a4[1], a4[2], a4[3]);
    *reinterpret_cast<float4*>(&Bs[lr][lc]) = make_float4(b4[0], b4[1], b4[2], b4[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {av4.x, av4.y, av4.z, av4.w};
      const float b[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* dW = p.dW + (long long)seg * p.m_real * p.n_real;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.m_real) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < p.n_real) atomicAdd(&dW[(long long)m * p.n_real + n], acc[i][j]);
    }
  }
}

// =============================================================================================
// Weight packing: fp32 Keras-layout master -> T [N][nseg*Cp + c] (K-major rows, zero padded).
// dst[n][k*Cp + c] = src[k*sk + n*sn + c*sc] for n < n_real, c < c_real.
// =============================================================================================
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ src, T* __restrict__ dst, int N, int n_real,
                                   int nseg, int Cp, int c_real, long long sk, long long sn, long long sc) {
  const long long total = (long long)N * nseg * Cp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const int k = (int)((i / Cp) % nseg);
    const int n = (int)(i / ((long long)Cp * nseg));
    float v = 0.f;
    if (n < n_real && c < c_real) v = src[k * sk + n * sn + c * sc];
    dst[i] = Elem<T>::from_f(v);
  }
}

// all packed copies of one model in ONE launch: blockIdx.y selects the op
// ld: destination row pitch (elements); op tap i < nseg reads source tap k0 + i*kstep and writes tap slot slot0 + i*slotstep
struct PackOp { const float* src; void* dst; int N, n_real, nseg, Cp, c_real; long long sk, sn, sc; long long ld; int k0, kstep, slot0, slotstep; };
struct PackOps { int n; PackOp op[24]; };
// grid (x = 32x32 tiles of the (n, c) plane, y = tap, z = op): reads follow the source's fastest axis, writes follow
// the destination's (c), transposing through shared memory when they differ.
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ PackOps ops) {
  __shared__ float tile[32][33];
  const PackOp& o = ops.op[blockIdx.z];
  if ((int)blockIdx.y >= o.nseg) return;
  const int k = o.k0 + (int)blockIdx.y * o.kstep, slot = o.slot0 + (int)blockIdx.y * o.slotstep;
  const int ct = (o.Cp + 31) / 32, ntl = (o.N + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  T* dst = reinterpret_cast<T*>(o.dst);
  for (int tIdx = blockIdx.x; tIdx < ct * ntl; tIdx += gridDim.x) {
    const int n0 = (tIdx / ct) * 32, c0 = (tIdx % ct) * 32;
    const bool n_fast = o.sn == 1;          // source contiguous along n (else along c or strided: read c-fastest)
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int a = ty + 8 * j;             // slow index of this read
      const int n = n_fast ? n0 + tx : n0 + a;
      const int c = n_fast ? c0 + a : c0 + tx;
      float v = 0.f;
      if (n < o.n_real && c < o.c_real) v = o.src[k * o.sk + n * o.sn + c * o.sc];
      tile[n - n0][c - c0] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty + 8 * j, c = c0 + tx;
      if (n < o.N && c < o.Cp) dst[(long long)n * o.ld + (long long)slot * o.Cp + c] = Elem<T>::from_f(tile[n - n0][c - c0]);
    }
  }
}

// =============================================================================================
// PhaseShuffle gather (calciumgan.py:117-138): X[b,t,:] = H[b, ps_index(t, shift[group(b)]), :].
// 16-byte vectors; shifts are per call-group (one scalar per layer per critic call).
// =============================================================================================
struct GroupShifts { int s[4]; };   // up to 3 groups used

template <typename T>
__global__ void ps_gather_kernel(const T* __restrict__ H, T* __restrict__ X, int Bt, int group_b, int w,
                                 int Cp, GroupShifts sh) {
  constexpr int V = 16 / sizeof(T);
  const int cv = Cp / V;
  const long long total = (long long)Bt * w * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const int t = (int)((i / cv) % w);
    const int b = (int)(i / ((long long)cv * w));
    const int j = ps_index(t, sh.s[b / group_b], w);
    const uint4 v = *reinterpret_cast<const uint4*>(H + ((long long)b * w + j) * Cp + c * V);
    *reinterpret_cast<uint4*>(X + ((long long)b * w + t) * Cp + c * V) = v;
  }
}

// Adjoint of the gather fused with the LeakyReLU slope of the layer below:
// DA[b,j,:] = slope(H[b,j,:]) * sum_{t : ps_index(t)=j} DX[b,t,:]      (at most 2 contributors)
template <typename T>
__global__ void ps_scatter_mask_kernel(const T* __restrict__ DX, const T* __restrict__ H, T* __restrict__ DA,
                                       int Bt, int group_b, int w, int Cp, GroupShifts sh) {
  pdl_enter();
  constexpr int V = Vec16<T>::N;
  const int cv = Cp / V;
  const long long total = (long long)Bt * w * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * V;
    const int j = (int)((i / cv) % w);
    const int b = (int)(i / ((long long)cv * w));
    const int s = sh.s[b / group_b];
    const T* dx = DX + (long long)b * w * Cp + c;
    float acc[V], t[V], h[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const int t0 = j - s;                 // direct: t + s = j
    if (t0 >= 0 && t0 < w) {
      vload<T>(dx + (long long)t0 * Cp, t);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] += t[e];
    }
    int t1 = -1;
    if (s > 0) {                          // reflected at the end: t + s > w-1, j = 2(w-1) - (t+s)
      t1 = 2 * (w - 1) - j - s;
      if (!(t1 >= 0 && t1 < w && t1 + s > w - 1)) t1 = -1;
    } else if (s < 0) {                   // reflected at the front: t + s < 0, j = -(t+s)
      t1 = -j - s;
      if (!(t1 >= 0 && t1 < w && t1 + s < 0)) t1 = -1;
    }
    if (t1 >= 0) {
      vload<T>(dx + (long long)t1 * Cp, t);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] += t[e];
    }
    const long long o = ((long long)b * w + j) * Cp + c;
    vload<T>(H + o, h);
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] *= lrelu_slope(h[e]);
    vstore<T>(DA + o, acc);
  }
}

// DA = DH * slope(H) (no PhaseShuffle / no layer-norm case)
template <typename T>
__global__ void mask_mul_kernel(const T* __restrict__ DH, const T* __restrict__ H, T* __restrict__ DA,
                                long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    DA[i] = Elem<T>::from_f(Elem<T>::to_f(DH[i]) * lrelu_slope(Elem<T>::to_f(H[i])));
}

// H = lrelu(A) (generator without layer-norm)
template <typename T>
__global__ void lrelu_kernel(const T* __restrict__ A, T* __restrict__ H, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    H[i] = Elem<T>::from_f(lrelu(Elem<T>::to_f(A[i])));
}

// =============================================================================================
// Critic input assembly + interpolation (wgan_gp.py:38-41): fp32 (B,L,C) -> T (B,L,Cp) groups.
// mode 0: X0[g0]=real, X0[g1]=fake, X0[g2]=alpha*real+(1-alpha)*fake ; mode 1: X0[g0]=src only.
// =============================================================================================
template <typename T>
__global__ void assemble_x0_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                   const float* __restrict__ alpha, T* __restrict__ X0, int B, int L, int C,
                                   int Cp, int mode) {
  pdl_enter();
  const int c4 = Cp / 4;
  const long long per = (long long)L * Cp;
  const long long total = (long long)B * L * c4;
  const long long gtot = (long long)B * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    const long long bt = i / c4;
    const int b = (int)(bt / L);
    float r[4] = {0.f, 0.f, 0.f, 0.f}, f[4] = {0.f, 0.f, 0.f, 0.f}, x[4] = {0.f, 0.f, 0.f, 0.f};
    const float a = mode == 0 ? alpha[b] : 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c + e < C) {
        r[e] = real[bt * C + c + e];
        if (mode == 0) {
          f[e] = fake[bt * C + c + e];
          x[e] = a * r[e] + (1.f - a) * f[e];
        }
      }
    }
    const long long o = bt * Cp + c;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      X0[o + e] = Elem<T>::from_f(r[e]);
      if (mode == 0) {
        X0[gtot + o + e] = Elem<T>::from_f(f[e]);
        X0[2 * gtot + o + e] = Elem<T>::from_f(x[e]);
      }
    }
  }
}

// xhat = alpha*real + (1-alpha)*fake (wgan_gp.py:38-41), fp32 in, T (B,L,Cp) out, 4 channels per thread
template <typename T>
__global__ void interp_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                              const float* __restrict__ alpha, T* __restrict__ Xh, int B, int L, int C, int Cp) {
  const int c4 = Cp / 4;
  const long long total = (long long)B * L * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    const long long bt = i / c4;
    const float a = alpha[(int)(bt / L)];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x = 0.f;
      if (c + e < C) x = a * real[bt * C + c + e] + (1.f - a) * fake[bt * C + c + e];
      Xh[bt * Cp + c + e] = Elem<T>::from_f(x);
    }
  }
}

// unpad + cast: T (B,L,Cp) -> fp32 (B,L,C)
template <typename T>
__global__ void unpad_kernel(const T* __restrict__ src, float* __restrict__ dst, long long rows, int C, int Cp) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    dst[i] = Elem<T>::to_f(src[(i / C) * Cp + c]);
  }
}

// =============================================================================================
// Critic head (calciumgan.py:188-190): scores[b] = <X5[b], wd> + bd ; one block per sample.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) head_forward_kernel(const T* __restrict__ X5, const float* __restrict__ wd,
                                                           const float* __restrict__ bd, float* __restrict__ scores,
                                                           int w5, int c5, int Cp) {
  pdl_enter();
  constexpr int V = Vec16<T>::N;
  __shared__ float red[8];
  const int b = blockIdx.x;
  const T* x = X5 + (long long)b * w5 * Cp;
  const int nv = w5 * Cp / V;
  float acc = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const int c = (i * V) % Cp, t = (i * V) / Cp;
    float v[V];
    vload<T>(x + (long long)i * V, v);
#pragma unroll
    for (int e = 0; e < V; ++e)
      if (c + e < c5) acc = fmaf(v[e], wd[t * c5 + c + e], acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) scores[b] = v + bd[0];
  }
}

// DA5[b,t,c] = slope(H5) * coef[b] * wd[t*c5+c]   (16-byte vectors)
template <typename T>
__global__ void head_backward_kernel(const T* __restrict__ H5, const float* __restrict__ wd,
                                     const float* __restrict__ coef, T* __restrict__ DA5, int Bt, int w5, int c5,
                                     int Cp) {
  pdl_enter();
  constexpr int V = Vec16<T>::N;
  const long long per = (long long)w5 * Cp;
  const long long total = (long long)Bt * per / V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long e0 = i * V;
    const int c = (int)(e0 % Cp);
    const int t = (int)((e0 / Cp) % w5);
    const int b = (int)(e0 / per);
    float h[V], o[V];
    vload<T>(H5 + e0, h);
    const float cb = coef[b];
#pragma unroll
    for (int e = 0; e < V; ++e) o[e] = c + e < c5 ? lrelu_slope(h[e]) * cb * wd[t * c5 + c + e] : 0.f;
    vstore<T>(DA5 + e0, o);
  }
}

// dwd[t*c5+c] += sum_b coef[b] * X5[b,t,c];  dbd += sum_{b<nb_bias} coef[b].  grid.y splits the batch; 16-byte loads.
// Samples b >= tail_from are read from X5_tail[b - tail_from] (the gradient penalty's linearised forward v_5, which
// must not overwrite the x_hat group's stored activation: its sign is that group's slope mask).
template <typename T>
__global__ void head_wgrad_kernel(const T* __restrict__ X5, const T* __restrict__ X5_tail, int tail_from,
                                  const float* __restrict__ coef, float* __restrict__ dwd,
                                  float* __restrict__ dbd, int Bt, int nb_bias, int w5, int c5, int Cp) {
  pdl_enter();
  constexpr int V = Vec16<T>::N;
  const int nv = w5 * Cp / V;
  const long long per_sample = (long long)w5 * Cp;
  const int per = (Bt + gridDim.y - 1) / gridDim.y;
  const int b_lo = blockIdx.y * per;
  const int b_hi = b_lo + per < Bt ? b_lo + per : Bt;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
    const int c = (i * V) % Cp, t = (i * V) / Cp;
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll 4
    for (int b = b_lo; b < b_hi; ++b) {
      float v[V];
      vload<T>((b < tail_from ? X5 + b * per_sample : X5_tail + (b - tail_from) * per_sample) + (long long)i * V, v);
      const float cb = coef[b];
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = fmaf(cb, v[e], acc[e]);
    }
    if (b_lo < b_hi) {
#pragma unroll
      for (int e = 0; e < V; ++e)
        if (c + e < c5) atomicAdd(&dwd[t * c5 + c + e], acc[e]);
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < nb_bias; ++b) s += coef[b];
    dbd[0] += s;
  }
}

// =============================================================================================
// column sums (bias gradients): out[c] += sum_rows X[row, c], c < c_real. grid-strided rows.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, float* __restrict__ out, long long rows,
                                                     int Cp, int c_real) {
  // blockDim = (64, 4): x over channels chunk, y over rows
  const int c = blockIdx.y * 64 + threadIdx.x;
  float acc = 0.f;
  if (c < c_real)
    for (long long r = (long long)blockIdx.x * 4 + threadIdx.y; r < rows; r += (long long)gridDim.x * 4)
      acc += Elem<T>::to_f(X[r * Cp + c]);
  __shared__ float red[4][64];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < c_real) {
    const float s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    atomicAdd(&out[c], s);
  }
}

// several column sums in one launch (blockIdx.y selects the tensor): bias gradients of all conv layers of a sub-step.
// 16-byte loads: thread = (row within the pass, 16-byte column vector), rows grid-strided; partial sums meet in shared
// memory, one atomicAdd per channel and block.
struct ColsumOps { int n; struct { const void* X; float* out; long long rows; int Cp, c_real; } op[8]; };
template <typename T>
__global__ void __launch_bounds__(256) colsum_multi_kernel(const __grid_constant__ ColsumOps ops) {
  pdl_enter();
  constexpr int V = Vec16<T>::N;
  const auto& o = ops.op[blockIdx.y];
  const int cv = o.Cp / V;                 // host: cv <= 256
  const int rpp = 256 / cv;                // rows per pass of this block
  const int tr = threadIdx.x / cv, tc = threadIdx.x - tr * cv;
  const T* X = reinterpret_cast<const T*>(o.X);
  float acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.f;
  if (tr < rpp) {
    const long long step = (long long)gridDim.x * rpp;
    long long r = (long long)blockIdx.x * rpp + tr;
    for (; r + 3 * step < o.rows; r += 4 * step) {   // four independent 16-byte loads in flight per thread
      float t0[V], t1[V], t2[V], t3[V];
      vload<T>(X + r * o.Cp + tc * V, t0);
      vload<T>(X + (r + step) * o.Cp + tc * V, t1);
      vload<T>(X + (r + 2 * step) * o.Cp + tc * V, t2);
      vload<T>(X + (r + 3 * step) * o.Cp + tc * V, t3);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] += (t0[e] + t1[e]) + (t2[e] + t3[e]);
    }
    for (; r < o.rows; r += step) {
      float t0[V];
      vload<T>(X + r * o.Cp + tc * V, t0);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] += t0[e];
    }
  }
  __shared__ float red[256 * V];
#pragma unroll
  for (int e = 0; e < V; ++e) red[threadIdx.x * V + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < o.c_real; c += 256) {
    float sum = 0.f;
    for (int k = 0; k < rpp; ++k) sum += red[(k * cv + c / V) * V + c % V];
    atomicAdd(&o.out[c], sum);
  }
}

// The same sums with a footprint that fits beside a resident tensor-core CTA (128 threads, no shared memory): partial sums
// go straight to the output with one atomicAdd per channel and thread. Runs on the engine's side stream under the GEMMs
// of the gradient-penalty passes; nobody reads the bias gradients before Adam.
template <typename T>
__global__ void __launch_bounds__(128) colsum_light_kernel(const __grid_constant__ ColsumOps ops) {
  constexpr int V = Vec16<T>::N;
  const auto& o = ops.op[blockIdx.y];
  const int cv = o.Cp / V;                 // host: cv <= 128
  const int rpp = 128 / cv;
  const int tr = threadIdx.x / cv, tc = threadIdx.x - tr * cv;
  if (tr >= rpp) return;
  const T* X = reinterpret_cast<const T*>(o.X);
  float acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.f;
  const long long step = (long long)gridDim.x * rpp;
  long long r = (long long)blockIdx.x * rpp + tr;
  for (; r + step < o.rows; r += 2 * step) {
    float t0[V], t1[V];
    vload<T>(X + r * o.Cp + tc * V, t0);
    vload<T>(X + (r + step) * o.Cp + tc * V, t1);
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] += t0[e] + t1[e];
  }
  for (; r < o.rows; r += step) {
    float t0[V];
    vload<T>(X + r * o.Cp + tc * V, t0);
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] += t0[e];
  }
#pragma unroll
  for (int e = 0; e < V; ++e)
    if (tc * V + e < o.c_real) atomicAdd(&o.out[tc * V + e], acc[e]);
}

// =============================================================================================
// Generator layer-norm + LeakyReLU (calciumgan.py:45-46): one warp per (b,t) row, channels C of Cp.
// =============================================================================================
template <typename T, int LPR>   // LPR lanes per row (8 | 16 | 32): short rows share a warp
__global__ void __launch_bounds__(256) ln_lrelu_forward_kernel(const T* __restrict__ A, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, T* __restrict__ H,
                                                               float* __restrict__ mu_out, float* __restrict__ rstd_out,
                                                               long long rows, int C, int Cp) {
  pdl_enter();
  constexpr int V = Vec16<T>::N;
  constexpr int MAXV = 4;               // vectors per lane: Cp <= LPR * 4 * V
  constexpr int RPW = 32 / LPR;         // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int nv = Cp / V;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const long long r = r0 + sub;
    const bool ok = r < rows;
    const T* a = A + (ok ? r : 0) * Cp;
    float x[MAXV][V];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int vi = sl + LPR * k;
      if (vi < nv) {
        vload<T>(a + vi * V, x[k]);
#pragma unroll
        for (int e = 0; e < V; ++e) s += x[k][e];   // pad channels hold exact zeros
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int vi = sl + LPR * k;
      if (vi < nv) {
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (vi * V + e < C) { const float d = x[k][e] - mu; q += d * d; }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + CG_LN_EPS);
    if (ok) {
      T* h = H + r * Cp;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int vi = sl + LPR * k;
        if (vi < nv) {
          float o[V];
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const int c = vi * V + e;
            o[e] = c < C ? lrelu((x[k][e] - mu) * rstd * gamma[c] + beta[c]) : 0.f;
          }
          vstore<T>(h + vi * V, o);
        }
      }
      if (sl == 0) {
        mu_out[r] = mu;
        rstd_out[r] = rstd;
      }
    }
  }
}

// backward of LN + LeakyReLU: DA, dgamma, dbeta. Same sub-warp row layout as the forward kernel; every lane owns fixed
// channel vectors, so dgamma / dbeta accumulate in registers over all rows of the thread and are reduced once per
// block through shared memory (the first version did two shared atomics per element).
template <typename T, int LPR, int MAXV>   // MAXV >= ceil((Cp / V) / LPR): channel vectors per lane (register arrays)
__global__ void __launch_bounds__(256, MAXV <= 2 ? 2 : 1) ln_lrelu_backward_kernel(
    const T* __restrict__ DH, const T* __restrict__ A, const T* __restrict__ H, const float* __restrict__ mu_in,
    const float* __restrict__ rstd_in, const float* __restrict__ gamma, T* __restrict__ DA,
    float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int C, int Cp) {
  pdl_enter();
  constexpr int V = Vec16<T>::N;
  constexpr int RPW = 32 / LPR;
  extern __shared__ float sm[];   // [2 * Cp]
  for (int i = threadIdx.x; i < 2 * Cp; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int nv = Cp / V;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float gacc[MAXV][V], bacc[MAXV][V], gam[MAXV][V];
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
#pragma unroll
    for (int e = 0; e < V; ++e) {
      gacc[k][e] = 0.f; bacc[k][e] = 0.f;
      const int c = (sl + LPR * k) * V + e;
      gam[k][e] = c < C ? gamma[c] : 0.f;
    }
  for (long long r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const long long r = r0 + sub;
    const bool ok = r < rows;
    const long long rr = ok ? r : 0;
    const float mu = mu_in[rr], rstd = rstd_in[rr];
    float xh[MAXV][V], dn[MAXV][V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int vi = sl + LPR * k;
      if (vi < nv) {
        float a[V], dh[V], h[V];
        vload<T>(A + rr * Cp + vi * V, a);
        vload<T>(DH + rr * Cp + vi * V, dh);
        vload<T>(H + rr * Cp + vi * V, h);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const bool real = vi * V + e < C && ok;
          xh[k][e] = real ? (a[e] - mu) * rstd : 0.f;
          dn[k][e] = real ? dh[e] * lrelu_slope(h[e]) : 0.f;
          const float gd = gam[k][e] * dn[k][e];
          s1 += gd;
          s2 += gd * xh[k][e];
          gacc[k][e] += dn[k][e] * xh[k][e];
          bacc[k][e] += dn[k][e];
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 /= C;
    s2 /= C;
    if (ok) {
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int vi = sl + LPR * k;
        if (vi < nv) {
          float o[V];
#pragma unroll
          for (int e = 0; e < V; ++e)
            o[e] = vi * V + e < C ? rstd * (gam[k][e] * dn[k][e] - s1 - xh[k][e] * s2) : 0.f;
          vstore<T>(DA + r * Cp + vi * V, o);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = sl + LPR * k;
    if (vi < nv) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        atomicAdd(&sm[vi * V + e], gacc[k][e]);
        atomicAdd(&sm[Cp + vi * V + e], bacc[k][e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dgamma[i], sm[i]);
    atomicAdd(&dbeta[i], sm[Cp + i]);
  }
}

// =============================================================================================
// Generator first Dense (calciumgan.py:32-34): HG0[b,t,c] = lrelu(z[b,:] . W0[:, t*nd+c] + b0), pad -> 0
// =============================================================================================
template <typename T>
__global__ void dense0_forward_kernel(const float* __restrict__ z, const float* __restrict__ W0,
                                      const float* __restrict__ b0, T* __restrict__ HG0, int B, int nd, int w0,
                                      int Cp) {
  pdl_enter();
  const long long total = (long long)B * w0 * Cp;
  const int nout = w0 * nd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const int t = (int)((i / Cp) % w0);
    const int b = (int)(i / ((long long)Cp * w0));
    float v = 0.f;
    if (c < nd) {
      const int j = t * nd + c;
      float acc = b0[j];
#pragma unroll 8
      for (int k = 0; k < nd; ++k) acc = fmaf(z[b * nd + k], W0[(long long)k * nout + j], acc);
      v = lrelu(acc);
    }
    HG0[i] = Elem<T>::from_f(v);
  }
}

// dW0[k][j] += sum_b z[b,k]*dpre[b,j] ; db0[j] += sum_b dpre[b,j] ; dpre = DHG0 * slope(HG0)
template <typename T>
__global__ void dense0_backward_kernel(const float* __restrict__ z, const T* __restrict__ DHG0,
                                       const T* __restrict__ HG0, float* __restrict__ dW0, float* __restrict__ db0,
                                       int B, int nd, int w0, int Cp) {
  pdl_enter();
  const int nout = w0 * nd;
  const int total = (nd + 1) * nout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % nout, k = i / nout;   // k == nd -> bias row
    const int t = j / nd, c = j % nd;
    float acc = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b) {
      const long long idx = ((long long)b * w0 + t) * Cp + c;
      const float dp = Elem<T>::to_f(DHG0[idx]) * lrelu_slope(Elem<T>::to_f(HG0[idx]));
      acc += (k < nd ? z[b * nd + k] : 1.f) * dp;
    }
    if (k < nd) dW0[(long long)k * nout + j] += acc;
    else db0[j] += acc;
  }
}

// DO[b,t,c] = DX0[b,t,c] * f(1-f) with f = fake32[b,t,c] (sigmoid head backward); pad -> 0.
// One 16-byte vector of DX0 / DO per thread; the unpadded fp32 rows are read as float2 (C even) or scalars.
template <typename T>
__global__ void sigmoid_backward_kernel(const T* __restrict__ DX0, const float* __restrict__ fake, T* __restrict__ DO,
                                        long long rows, int C, int Cp, int normalize) {
  pdl_enter();
  constexpr int V = Vec16<T>::N;
  const int cv = Cp / V;
  const long long total = rows * cv;
  const bool pairs = (C & 1) == 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cv;
    const int c = (int)(i - row * cv) * V;
    float v[V];
    if (c < C) {
      vload<T>(DX0 + row * Cp + c, v);
      const float* f = fake + row * C + c;
#pragma unroll
      for (int e = 0; e < V; e += 2) {
        float f0 = 0.f, f1 = 0.f;
        if (normalize) {
          if (pairs) { if (c + e < C) { const float2 t = *reinterpret_cast<const float2*>(f + e); f0 = t.x; f1 = t.y; } }
          else { if (c + e < C) f0 = f[e]; if (c + e + 1 < C) f1 = f[e + 1]; }
          v[e] *= f0 * (1.f - f0);
          v[e + 1] *= f1 * (1.f - f1);
        }
        if (c + e >= C) v[e] = 0.f;
        if (c + e + 1 >= C) v[e + 1] = 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] = 0.f;
    }
    vstore<T>(DO + row * Cp + c, v);
  }
}

// =============================================================================================
// Gradient-penalty scalars (wgan_gp.py:49-50,58-62)
// =============================================================================================
// sumsq[b] = sum g[b,:]^2 ; one block per (sample, chunk), atomics into sumsq (zeroed before). The accumulator is a double:
// adding a few hundred fp32 partial sums of similar magnitude in double is exact, hence independent of the arrival order
template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ G, double* __restrict__ sumsq,
                                                    long long per_sample, int chunks) {
  __shared__ float red[8];
  const int b = blockIdx.x / chunks, ch = blockIdx.x % chunks;
  const long long len = (per_sample + chunks - 1) / chunks;
  const long long beg = (long long)ch * len;
  long long end = beg + len;
  if (end > per_sample) end = per_sample;
  const T* g = G + (long long)b * per_sample;
  float acc = 0.f;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const float v = Elem<T>::to_f(g[i]);
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(&sumsq[b], (double)v);
  }
}
__global__ void sumsq_to_float_kernel(const double* __restrict__ sumsq, float* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = (float)sumsq[b];
}

// single block: losses + per-sample coefficient of u = d(lambda*GP)/dg
// scal[0]=dis_loss scal[1]=gp scal[2]=real_loss scal[3]=fake_loss ; norms[b] = ||g_b||
__global__ void critic_scalars_kernel(const float* __restrict__ scores, const double* __restrict__ sumsq,
                                      float* __restrict__ ucoef, float* __restrict__ norms, float* __restrict__ scal,
                                      int B, float lambda) {
  pdl_enter();
  __shared__ float red[3][32];
  float sr = 0.f, sf = 0.f, sg = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    sr += scores[b];
    sf += scores[B + b];
    const float n = sqrtf((float)sumsq[b]);
    norms[b] = n;
    sg += (n - 1.f) * (n - 1.f);
    ucoef[b] = lambda * (2.f / B) * (n - 1.f) / n;   // no epsilon: tf.norm (wgan_gp.py:49)
  }
  sr = warp_sum(sr); sf = warp_sum(sf); sg = warp_sum(sg);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = sr; red[1][threadIdx.x >> 5] = sf; red[2][threadIdx.x >> 5] = sg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b2 = 0.f, c = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { a += red[0][i]; b2 += red[1][i]; c += red[2][i]; }
    const float real_loss = -a / B, fake_loss = b2 / B, gp = c / B;
    scal[0] = real_loss + fake_loss + lambda * gp;
    scal[1] = gp;
    scal[2] = real_loss;
    scal[3] = fake_loss;
  }
}

// gen_loss = -mean(scores[0:B]) -> scal[4]
__global__ void gen_loss_kernel(const float* __restrict__ scores, float* __restrict__ scal, int B) {
  pdl_enter();
  __shared__ float red[32];
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s += scores[b];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) a += red[i];
    scal[4] = -a / B;
  }
}

// coef for the concatenated critic batch: [-1/B]*B, [+1/B]*B, [1]*B   (or a constant for 1 group)
__global__ void fill_coef_kernel(float* coef, int B, int groups, float single) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * groups) return;
  if (groups == 1) coef[i] = single;
  else coef[i] = i < B ? -1.f / B : (i < 2 * B ? 1.f / B : 1.f);
}

// V0[b] = ucoef[b] * G[b]   (16-byte vectors; per_sample % (16/sizeof(T)) == 0)
template <typename T>
__global__ void scale_rows_kernel(const T* __restrict__ G, const float* __restrict__ ucoef, T* __restrict__ V,
                                  long long per_sample, long long total) {
  pdl_enter();
  constexpr int W = Vec16<T>::N;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * W; i < total;
       i += (long long)gridDim.x * blockDim.x * W) {
    float v[W];
    vload<T>(G + i, v);
    const float u = ucoef[i / per_sample];
#pragma unroll
    for (int e = 0; e < W; ++e) v[e] *= u;
    vstore<T>(V + i, v);
  }
}

// =============================================================================================
// Signal metrics (gan.py:32-41, signals_metrics.py:9-28): warp per (b,t) row over C neurons.
// acc[0..3] += squared diff of min / max / mean / std (population). Finalised by /rows on host side kernel.
// =============================================================================================
__global__ void __launch_bounds__(256) metrics_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                                      float* __restrict__ acc, long long rows, int C, float smin,
                                                      float smax, int normalize) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float sc = normalize ? (smax - smin) : 1.f, of = normalize ? smin : 0.f;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long r = warp; r < rows; r += nwarps) {
    float st[2][4];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const float* x = (w == 0 ? real : fake) + r * C;
      float mn = INFINITY, mx = -INFINITY, s = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float v = x[c] * sc + of;
        mn = fminf(mn, v); mx = fmaxf(mx, v); s += v;
      }
      mn = warp_min(mn); mx = warp_max(mx);
      const float mean = warp_sum(s) / C;
      float q = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float d = x[c] * sc + of - mean;
        q += d * d;
      }
      st[w][0] = mn; st[w][1] = mx; st[w][2] = mean; st[w][3] = sqrtf(warp_sum(q) / C);
    }
    a0 += (st[0][0] - st[1][0]) * (st[0][0] - st[1][0]);
    a1 += (st[0][1] - st[1][1]) * (st[0][1] - st[1][1]);
    a2 += (st[0][2] - st[1][2]) * (st[0][2] - st[1][2]);
    a3 += (st[0][3] - st[1][3]) * (st[0][3] - st[1][3]);
  }
  if (lane == 0) {
    const float inv = 1.f / (float)rows;
    atomicAdd(&acc[0], a0 * inv); atomicAdd(&acc[1], a1 * inv);
    atomicAdd(&acc[2], a2 * inv); atomicAdd(&acc[3], a3 * inv);
  }
}

// Same statistics with 8 lanes per (b,t) row (four rows per warp) for even C <= 128: each lane keeps its <= 16
// elements of the real and of the fake row in registers (float2 loads, all issued up front), so the exact two-pass
// variance costs no second read and every reduction is 3 shuffles instead of 5 per row.
__global__ void __launch_bounds__(256) metrics8_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                                       float* __restrict__ acc, long long rows, int C, float smin,
                                                       float smax, int normalize) {
  const int sub = threadIdx.x & 7;
  const long long grp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long ngrp = ((long long)gridDim.x * blockDim.x) >> 3;
  const float sc = normalize ? (smax - smin) : 1.f, of = normalize ? smin : 0.f;
  const int half = C >> 1;   // float2 per row
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long r0 = 0; r0 < rows; r0 += ngrp) {   // every lane runs the same number of iterations (full-mask shuffles)
    const long long r = r0 + grp;
    const bool rok = r < rows;
    float2 v[2][8];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const float2* x = reinterpret_cast<const float2*>((w == 0 ? real : fake) + (rok ? r : 0) * C);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[w][k] = (rok && sub + 8 * k < half) ? x[sub + 8 * k] : make_float2(0.f, 0.f);
    }
    float st[2][4];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      float mn = INFINITY, mx = -INFINITY, s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (sub + 8 * k < half) {
          v[w][k].x = v[w][k].x * sc + of; v[w][k].y = v[w][k].y * sc + of;
          mn = fminf(mn, fminf(v[w][k].x, v[w][k].y)); mx = fmaxf(mx, fmaxf(v[w][k].x, v[w][k].y));
          s += v[w][k].x + v[w][k].y;
        }
      }
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        s += __shfl_xor_sync(0xffffffffu, s, o);
      }
      const float mean = s / C;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (sub + 8 * k < half) {
          const float d0 = v[w][k].x - mean, d1 = v[w][k].y - mean;
          q += d0 * d0 + d1 * d1;
        }
      }
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      st[w][0] = mn; st[w][1] = mx; st[w][2] = mean; st[w][3] = sqrtf(q / C);
    }
    if (rok && sub == 0) {
      a0 += (st[0][0] - st[1][0]) * (st[0][0] - st[1][0]);
      a1 += (st[0][1] - st[1][1]) * (st[0][1] - st[1][1]);
      a2 += (st[0][2] - st[1][2]) * (st[0][2] - st[1][2]);
      a3 += (st[0][3] - st[1][3]) * (st[0][3] - st[1][3]);
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  if ((threadIdx.x & 31) == 0) {
    const float inv = 1.f / (float)rows;
    atomicAdd(&acc[0], a0 * inv); atomicAdd(&acc[1], a1 * inv);
    atomicAdd(&acc[2], a2 * inv); atomicAdd(&acc[3], a3 * inv);
  }
}

// Data-parallel gradient exchange over NVLink peer memory (no reference counterpart): every rank's flat
// gradient buffer is mapped into every other rank (symmetric memory); this kernel pulls the W buffers with 16-byte
// peer loads and writes their sum, always added in rank order 0 .. W-1 so that all ranks produce bit-identical results.
// Footprint by design: 64 threads (<= 128 registers each = 8 K of the ~11.7 K registers a resident tcgen05 GEMM CTA leaves
// free), no shared memory -- a CTA of it fits beside a GEMM CTA, so the exchange runs UNDER the GEMMs of the next
// sub-step's generator forward instead of waiting for an SM to drain.
struct PeerPtrs { const float* p[8]; };
// out[first .. first + n) = sum_r peers[r][first .. first + n)   (first % 4 == 0). Few CTAs, U vectors per peer in flight
// per thread: the kernel shares every SM it lands on with a GEMM CTA, so it is sized to disturb few of them.
template <int W, int U>
__global__ void __launch_bounds__(64) peer_sum_kernel(const __grid_constant__ PeerPtrs peers, float* __restrict__ out_all,
                                                       long long first, long long n) {
  PeerPtrs pp = peers;
#pragma unroll
  for (int r = 0; r < W; ++r) pp.p[r] += first;
  float* out = out_all + first;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
#pragma unroll
        for (int r = 0; r < W; ++r)
          asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[u][r].x), "=f"(v[u][r].y), "=f"(v[u][r].z), "=f"(v[u][r].w)
                       : "l"(reinterpret_cast<const float4*>(pp.p[r]) + i));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        float4 a = v[u][0];
#pragma unroll
        for (int r = 1; r < W; ++r) { a.x += v[u][r].x; a.y += v[u][r].y; a.z += v[u][r].z; a.w += v[u][r].w; }
        reinterpret_cast<float4*>(out)[i] = a;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < W; ++r) a += pp.p[r][i];
    out[i] = a;
  }
}
// second phase of the two-phase exchange: slice s of the reduced gradient was summed by rank s; every rank copies the
// W - 1 slices it does not own from their owners' (peer-mapped) reduced buffers into its own. blockIdx.y = slice.
__global__ void __launch_bounds__(64) peer_gather_kernel(const __grid_constant__ PeerPtrs peers, float* __restrict__ out,
                                                          long long slice, long long n, int rank) {
  const int s = blockIdx.y;
  if (s == rank) return;
  const long long first = (long long)s * slice;
  long long cnt = n - first;
  if (cnt <= 0) return;
  if (cnt > slice) cnt = slice;
  const float* src = peers.p[s] + first;
  float* dst = out + first;
  const long long n4 = cnt >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i0 + u * stride < n4)
        asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                     : "l"(reinterpret_cast<const float4*>(src) + i0 + u * stride));
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i0 + u * stride < n4) reinterpret_cast<float4*>(dst)[i0 + u * stride] = v[u];
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(cnt & 3)) dst[(n4 << 2) + threadIdx.x] = src[(n4 << 2) + threadIdx.x];
}

// dst[i, :] = src[idx[i], :] for a device-resident dataset cache (the reference caches its dataset in host memory,
// gan/utils/dataset_helper.py:171 `train_ds.cache()`, and shuffles indices): one block column per output row, 16-byte
// copies; a step then moves a few hundred indices over PCIe instead of the batch.
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx,
                                                          float* __restrict__ dst, long long row_elems, long long n_src) {
  const long long i = blockIdx.y;
  long long r = idx[i];
  if (r < 0 || r >= n_src) r = 0;   // indices are validated on the host; never read out of bounds
  const float4* s4 = reinterpret_cast<const float4*>(src + r * row_elems);
  float4* d4 = reinterpret_cast<float4*>(dst + i * row_elems);
  const long long n4 = row_elems >> 2;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (long long)gridDim.x * blockDim.x) d4[j] = s4[j];
  if (blockIdx.x == 0 && threadIdx.x < (row_elems & 3))
    dst[i * row_elems + (n4 << 2) + threadIdx.x] = src[r * row_elems + (n4 << 2) + threadIdx.x];
}

// Same statistics, staged through shared memory: a block copies 32 consecutive rows of both tensors (32 * C floats, a
// multiple of 16 bytes for every C) with fully coalesced 16-byte loads, then 8 lanes per row reduce from shared memory.
// The unpadded rows are 408 bytes at C = 102, so per-row global loads can be at most 8 bytes wide and leave half of
// every 32-byte sector request unused; this form reads each byte once at full width (measured 102 -> ~45 us per step).
__global__ void __launch_bounds__(256) metrics_smem_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                                           float* __restrict__ acc, long long rows, int C, float smin,
                                                           float smax, int normalize) {
  extern __shared__ __align__(16) float msm[];   // [2][32 * C]
  const float sc = normalize ? (smax - smin) : 1.f, of = normalize ? smin : 0.f;
  const int sub = threadIdx.x & 7, rloc = threadIdx.x >> 3;   // 32 rows x 8 lanes
  const int chunk_f = 32 * C, chunk_v = chunk_f >> 2;
  const long long nchunks = (rows + 31) >> 5;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long r0 = ch << 5;
    const int nrow = rows - r0 < 32 ? (int)(rows - r0) : 32;
    const int nv = (nrow * C) >> 2, tail = (nrow * C) & 3;
    __syncthreads();   // the previous chunk has been consumed
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const float* src = (w == 0 ? real : fake) + r0 * C;
      float4* dst = reinterpret_cast<float4*>(msm + w * chunk_f);
      for (int i = threadIdx.x; i < nv; i += 256) dst[i] = reinterpret_cast<const float4*>(src)[i];
      if (threadIdx.x < tail) msm[w * chunk_f + (nv << 2) + threadIdx.x] = src[(nv << 2) + threadIdx.x];
    }
    (void)chunk_v;
    __syncthreads();
    float st[2][4];
    const bool rok = rloc < nrow;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const float* x = msm + w * chunk_f + rloc * C;
      float mn = INFINITY, mx = -INFINITY, s = 0.f;
      if (rok)
        for (int c = sub; c < C; c += 8) {
          const float v = x[c] * sc + of;
          mn = fminf(mn, v); mx = fmaxf(mx, v); s += v;
        }
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        s += __shfl_xor_sync(0xffffffffu, s, o);
      }
      const float mean = s / C;
      float q = 0.f;
      if (rok)
        for (int c = sub; c < C; c += 8) {
          const float d = x[c] * sc + of - mean;
          q = fmaf(d, d, q);
        }
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      st[w][0] = mn; st[w][1] = mx; st[w][2] = mean; st[w][3] = sqrtf(q / C);
    }
    if (rok && sub == 0) {
      a0 += (st[0][0] - st[1][0]) * (st[0][0] - st[1][0]);
      a1 += (st[0][1] - st[1][1]) * (st[0][1] - st[1][1]);
      a2 += (st[0][2] - st[1][2]) * (st[0][2] - st[1][2]);
      a3 += (st[0][3] - st[1][3]) * (st[0][3] - st[1][3]);
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  __shared__ float red[8][4];
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = a0; red[threadIdx.x >> 5][1] = a1; red[threadIdx.x >> 5][2] = a2; red[threadIdx.x >> 5][3] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(&acc[threadIdx.x], t / (float)rows);
  }
}

// out = x*(max-min)+min (gan/utils/utils.py:30-32)
__global__ void denorm_kernel(const float* __restrict__ x, float* __restrict__ out, long long total, float smin,
                              float smax) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = x[i] * (smax - smin) + smin;
}

// =============================================================================================
// Adam (Keras form, optimizer.py:9,34): fp32 master weights + moments; w -= lr_t * m / (sqrt(v) + eps_hat),
// lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t). The iteration counter lives on the device so that a step whose gradient
// holds a non-finite value can be skipped (what the reference's LossScaleOptimizer does, optimizer.py:10-12) without
// a host round trip, and so that no kernel argument changes from step to step.
// =============================================================================================
struct OptState {
  long long steps;        // optimizer.iterations
  long long skipped;      // updates skipped because of a non-finite gradient
  float lr_t;             // bias-corrected step size of the update in flight
  int do_update;          // 0: the update in flight is skipped
  unsigned int bad;       // scratch: a non-finite gradient was seen
  unsigned int done;      // scratch: blocks of adam_prepare_kernel that have finished
};

// Pass 1 over the gradient: any non-finite value? The last block to finish advances the counters and computes lr_t.
__global__ void __launch_bounds__(256) adam_prepare_kernel(const float* __restrict__ g, long long n, OptState* st, float lr,
                                                           float b1, float b2) {
  pdl_enter();
  unsigned int bad = 0;
  const long long n4 = n >> 2;
  const uint4* g4 = reinterpret_cast<const uint4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = g4[i];
    bad |= ((v.x & 0x7f800000u) == 0x7f800000u) | ((v.y & 0x7f800000u) == 0x7f800000u) |
           ((v.z & 0x7f800000u) == 0x7f800000u) | ((v.w & 0x7f800000u) == 0x7f800000u);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3))
    bad |= (__float_as_uint(g[(n4 << 2) + threadIdx.x]) & 0x7f800000u) == 0x7f800000u;
  bad = __syncthreads_or((int)bad);
  __shared__ unsigned int ticket;
  if (threadIdx.x == 0) {
    if (bad) atomicOr(&st->bad, 1u);
    __threadfence();
    ticket = atomicAdd(&st->done, 1u);
  }
  __syncthreads();
  if (ticket == gridDim.x - 1 && threadIdx.x == 0) {
    __threadfence();
    const unsigned int any_bad = atomicOr(&st->bad, 0u);
    if (any_bad) {
      st->skipped += 1;
      st->do_update = 0;
    } else {
      st->steps += 1;
      const double t = (double)st->steps;
      st->lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
      st->do_update = 1;
    }
    st->bad = 0u;
    st->done = 0u;
  }
}

__device__ __forceinline__ void adam_elem(float& w, float& m, float& v, float g, float lr_t, float b1, float b2, float eps) {
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  w -= lr_t * m / (sqrtf(v) + eps);
}

__global__ void adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ g, long long n, const OptState* __restrict__ st, float b1,
                            float b2, float eps, float gscale) {
  if (!st->do_update) return;
  const float lr_t = st->lr_t;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float wi = w[i], mi = m[i], vi = v[i];
    adam_elem(wi, mi, vi, g[i] * gscale, lr_t, b1, b2, eps);
    m[i] = mi; v[i] = vi; w[i] = wi;
  }
}

// Adam and the refresh of the packed low-precision GEMM operands in ONE pass over the parameters. GEMM kernels are
// walked as 32 x 32 tiles of one tap's (a, b) plane (b = the Keras layout's fastest axis): w, m, v, g are read and
// written coalesced along b, the packed copy whose rows run along b is written from registers, the one whose rows run
// along a through a shared-memory transpose. Everything else (biases, layer-norm parameters, the two small dense
// layers) is a plain element-wise range. Padded rows / columns of the packed copies are never touched (they are zero
// from cg_create / the last full re-pack and no update changes them).
struct AdamTensor {
  long long off;          // offset in the flat parameter arrays
  int K, A, B;            // taps, slow and fast extent of one tap's plane: element (k, a, b) at off + (k*A + a)*B + b
  int Ap, Bp;             // padded extents in the packed copies
  void* direct;           // [a][k*Bp + b]
  void* trans;            // [b][k*Ap + a]
  void* trans2;           // optional row-pair copy [g*Bp + b][(k + 2g)*Ap + a], g = 0, 1, row pitch (K + 2)*Ap (see RsParams.row_pairs)
  void* direct2;          // optional merged-phase copy [g*Ap + a][j*Bp + b]: tap k belongs to output phase g and input window j
  int d2_emin, d2_K2;     //   (RsParams.merged_phases): d = padL - k, g = d & 1, j = (d + g)/2 - emin, row pitch K2*Bp
  int tiles_a, tiles_b;
  long long item0;        // first work item of this tensor
};
struct AdamRange { long long off, len, item0; };
struct AdamPlan {
  int nt, nr;
  AdamTensor t[8];
  AdamRange r[16];
  long long items;        // tile items first, then 1024-element chunks of the ranges
  long long tile_items;
};

template <typename T>
__global__ void __launch_bounds__(256) adam_pack_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                                                        const float* __restrict__ g, const __grid_constant__ AdamPlan plan,
                                                        const OptState* __restrict__ st, float b1, float b2, float eps,
                                                        float gscale) {
  pdl_enter();
  if (!st->do_update) return;
  const float lr_t = st->lr_t;
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (long long it = blockIdx.x; it < plan.items; it += gridDim.x) {
    if (it < plan.tile_items) {
      int ti = 0;
#pragma unroll
      for (int j = 1; j < 8; ++j) if (j < plan.nt && it >= plan.t[j].item0) ti = j;
      const AdamTensor& t = plan.t[ti];
      int r = (int)(it - t.item0);
      const int tb = r % t.tiles_b; r /= t.tiles_b;
      const int ta = r % t.tiles_a;
      const int k = r / t.tiles_a;
      const int a0 = ta * 32, b = tb * 32 + tx;
      T* direct = reinterpret_cast<T*>(t.direct);
      T* trans = reinterpret_cast<T*>(t.trans);
      __syncthreads();   // the previous item's transposed reads are done
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = a0 + ty + 8 * j;
        float wi = 0.f;
        if (a < t.A && b < t.B) {
          const long long i = t.off + ((long long)k * t.A + a) * t.B + b;
          wi = w[i];
          float mi = m[i], vi = v[i];
          adam_elem(wi, mi, vi, g[i] * gscale, lr_t, b1, b2, eps);
          m[i] = mi; v[i] = vi; w[i] = wi;
          direct[((long long)a * t.K + k) * t.Bp + b] = Elem<T>::from_f(wi);
          if (t.direct2) {
            const int d = (t.K - 2) / 2 - k, g2 = d & 1, j = ((d + g2) >> 1) - t.d2_emin;
            reinterpret_cast<T*>(t.direct2)[((long long)(g2 * t.Ap + a) * t.d2_K2 + j) * t.Bp + b] = Elem<T>::from_f(wi);
          }
        }
        tile[ty + 8 * j][tx] = wi;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int bb = tb * 32 + ty + 8 * j, a = a0 + tx;
        if (bb < t.B && a < t.A) {
          const T val = Elem<T>::from_f(tile[tx][ty + 8 * j]);
          trans[((long long)bb * t.K + k) * t.Ap + a] = val;
          if (t.trans2) {
            T* t2 = reinterpret_cast<T*>(t.trans2);
            const long long ld2 = (long long)(t.K + 2) * t.Ap;
            t2[(long long)bb * ld2 + (long long)k * t.Ap + a] = val;
            t2[(long long)(t.Bp + bb) * ld2 + (long long)(k + 2) * t.Ap + a] = val;
          }
        }
      }
    } else {
      const long long e = it - plan.tile_items;
      int ri = 0;
#pragma unroll
      for (int j = 1; j < 16; ++j) if (j < plan.nr && e >= plan.r[j].item0) ri = j;
      const AdamRange& rg = plan.r[ri];
      const long long base = (e - rg.item0) * 1024;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long o = base + j * 256 + threadIdx.x;
        if (o < rg.len) {
          const long long i = rg.off + o;
          float wi = w[i], mi = m[i], vi = v[i];
          adam_elem(wi, mi, vi, g[i] * gscale, lr_t, b1, b2, eps);
          m[i] = mi; v[i] = vi; w[i] = wi;
        }
      }
    }
  }
}

// =============================================================================================
// Philox4x32-10 counter RNG for noise (gan.py:29-30) and interpolation alpha (wgan_gp.py:40).
// =============================================================================================
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t stream, uint64_t idx, uint32_t (&c)[4]) {
  c[0] = (uint32_t)idx; c[1] = (uint32_t)(idx >> 32); c[2] = (uint32_t)stream; c[3] = (uint32_t)(stream >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// mode 0: standard normal (Box-Muller); mode 1: uniform [0,1). blockIdx.y = draw number: out + y * n, stream0 + (y << 24)
__global__ void rng_fill_kernel(float* __restrict__ out_all, long long n, uint64_t seed, uint64_t stream0, int mode) {
  float* out = out_all + (long long)blockIdx.y * n;
  const uint64_t stream = stream0 + ((uint64_t)blockIdx.y << 24);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n;
       i += (long long)gridDim.x * blockDim.x) {
    uint32_t c[4];
    philox4x32(seed, stream, (uint64_t)i, c);
    float u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = (c[j] >> 8) * (1.0f / 16777216.0f);   // [0,1)
    float r[4];
    if (mode == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = u[j];
    } else {
      const float ra = sqrtf(-2.f * logf(1.f - u[0])), rb = sqrtf(-2.f * logf(1.f - u[2]));
      r[0] = ra * cospif(2.f * u[1]); r[1] = ra * sinpif(2.f * u[1]);
      r[2] = rb * cospif(2.f * u[3]); r[3] = rb * sinpif(2.f * u[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i * 4 + j < n) out[i * 4 + j] = r[j];
  }
}
