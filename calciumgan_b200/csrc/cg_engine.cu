// calciumgan_b200 engine: context, device arena, step orchestration and the C ABI
// declared in include/calciumgan_b200.h.  Replaces TensorFlow underneath the reference's
// gan/algorithms/wgan_gp.py + gan/models/calciumgan.py (paths relative to /root/reference).
//
// Data layout in HBM (one arena per context):
//   activations   channels-last (batch, time, Cp), Cp = channels rounded up to 64, pad == 0
//   critic batch  the three critic calls of one sub-step (real, fake, xhat) are ONE batch of 3B
//                 samples ("groups"), each group with its own 4 PhaseShuffle shifts
//   weights       fp32 master in Keras get_weights() order (flat), Adam m/v alongside, plus
//                 packed T copies [N][tap*Cp + c] for the forward and data-gradient GEMMs
//   gradients     flat fp32 in the same order as the master weights (all-reduced in place by DP)
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/calciumgan_b200.h"
#include "cg_kernels_simt.cuh"
#include "cg_kernels_tc.cuh"
#include "cg_kernels_head.cuh"

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int set_err(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
int cg_tc_set_err(const char* msg) { return set_err("%s", msg); }
#define CU(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess)                                                                 \
      return set_err("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_));    \
  } while (0)
#define CK(x)              \
  do {                     \
    int r_ = (x);          \
    if (r_) return r_;     \
  } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static const int NL = 5;   // conv layers per network (calciumgan.py: 5 x Conv1DTranspose / 5 x Conv1D)

struct ParamInfo {
  int64_t shape[4];
  int ndim;
  int64_t offset, size;
};

struct Model {
  std::vector<ParamInfo> params;
  int64_t total = 0;
  float *w = nullptr, *m = nullptr, *v = nullptr, *g = nullptr;   // flat fp32
  OptState* opt = nullptr;                                         // device: optimizer.iterations, skipped updates, lr_t
  float* g_own = nullptr;      // the library's own gradient buffer (g may point at a caller-provided symmetric buffer)
  float* gr = nullptr;         // data parallel over peer memory: sum of every rank's gradients (cg_reduce_peer_grads)
  void add(std::initializer_list<int64_t> shp) {
    ParamInfo p{};
    p.ndim = (int)shp.size();
    int i = 0;
    p.size = 1;
    for (auto s : shp) { p.shape[i++] = s; p.size *= s; }
    p.offset = total;
    total += p.size;
    params.push_back(p);
  }
};

struct cg_ctx {
  cg_config cfg;
  bool bf;          // bf16 activations
  bool use_tc;      // tcgen05 implicit-GEMM kernels
  int esz;
  cudaStream_t stream = 0;
  int64_t launches = 0;
  int64_t tc_launches = 0;
  bool profiling = false;
  struct ProfRec { cudaEvent_t e0, e1; int cls; double flops; double bytes; char desc[96]; };   // cls 3: memory-bound glue
  std::vector<ProfRec> prof;
  bool glue_open = false;                 // a glue record waits for its launch (closed by post_launch)
  int64_t dev_bytes = 0;
  std::vector<void*> allocs;

  int L, C, nd, nu, K, w0, Bmax;
  int gc[NL + 1], gcp[NL + 1], gl[NL + 1];   // generator channels / padded / lengths
  int dc[NL + 1], dcp[NL + 1], dl[NL + 1];   // critic
  Model gen, dis;
  int g_k[NL + 1], g_b[NL + 1], g_gam[NL + 1], g_bet[NL + 1], g_d1k, g_d1b;   // param indices
  // packed weights (T)
  void *Wf_d[NL + 1], *Wb_d[NL + 1];     // critic conv: forward [Cout][K*Cin_p], dgrad [Cin][K*Cout_p]
  void *Wf_g[NL + 1], *Wb_g[NL + 1];     // generator convT: forward [Cout][K*Cin_p], bwd-data [Cin][K*Cout_p]
  void *Wf_d1, *Wb_d1;                   // generator output dense: [C][Cp], transposed
  void* Wf2_d[NL + 1];                   // critic conv, row-pair form (64 output channels): [2*64][(K+2)*Cin_p], else null
  void* Wb2_d[NL + 1];                   // critic conv data gradient, merged output phases (64 input channels): [2*64][K2*Cout_p], else null
  int mp_emin = 0, mp_K2 = 0;            // merged phases: first input-window shift and number of windows (K = 24: -6, 13)
  // critic activations (capacity 3*Bmax)
  void *X[NL + 1], *H[NL + 1], *DX[NL + 1], *DA[NL + 1];
  void* V5;                                // gradient penalty: linearised forward output of the last conv layer (Bmax samples)
  float *scores, *coef, *ucoef, *norms;
  double* sumsq;   // per-sample ||g||^2, accumulated in double: the sum of the fp32 partials is then exact, so the result
                   // does not depend on the order in which the atomics of different CTAs arrive
  // generator activations (capacity Bmax)
  float* Z;
  void *HG[NL + 1], *AG[NL + 1], *DHG[NL + 1], *DAG[NL + 1], *DO;
  float *MU[NL + 1], *RSTD[NL + 1];
  float* FAKE32;
  float *alpha_buf, *noise_buf;
  float* d_scal;        // [n_critic+1][CG_NUM_SCALARS]
  float* h_scal;        // pinned
  // Random streams (all functions of (seed, number of step calls so far), never of what the caller injected):
  // noise / alpha are rank-specific Philox streams, PhaseShuffle shifts one rank-SHARED stream (per-call scalars,
  // calciumgan.py:121-124) -- every step function advances each counter by a fixed amount whether or not it used it
  uint64_t seed = 1234, noise_calls = 0, alpha_calls = 0, shift_draws = 0;
  // draws of the last step function (cg_debug_last_draws): device pointers into caller / library buffers, host shifts
  const float *last_noise = nullptr, *last_alpha = nullptr;
  int64_t last_n_noise = 0, last_n_alpha = 0;
  std::vector<int32_t> last_shifts;
  int dbg_flags = 0;                       // CG_DEBUG_* (cfg.debug_flags | environment, fixed at cg_create)
  // side stream for memory-bound work that has no consumer until the end of the step (signal metrics): a 128-thread
  // block of it fits beside a resident tensor-core CTA, so it runs UNDER the generator's backward GEMMs
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // cg_prefetch_generator: the generator forward of the next sub-step has already run (data parallel overlap)
  struct { bool valid = false; int for_gen_step = 0, B = 0; const float* noise = nullptr; const float* alpha = nullptr; } pref;
  AdamPlan adam_plan[2];
  static const int NB = 3;                 // gradient buckets per model (ready order: last layers first)
  cudaEvent_t bucket_evt[2][NB] = {};
  int64_t bucket_off[2][NB + 1] = {};
  TcState tc;
};

// ------------------------------------------------------------------------------------------ arena
static int dalloc(cg_ctx* c, void** p, size_t bytes) {
  bytes = (bytes + 1023) / 1024 * 1024;
  CU(cudaMalloc(p, bytes));
  CU(cudaMemsetAsync(*p, 0, bytes, c->stream));
  c->allocs.push_back(*p);
  c->dev_bytes += (int64_t)bytes;
  return 0;
}

// ------------------------------------------------------------------------------------------ seg tables
static int floor_div2(int d) { return d >= 0 ? d / 2 : -((-d + 1) / 2); }

// strided-conv form: A viewed as (B, Lin/2, 2*Cp); tap k -> row shift j, column offset p*Cp
static SegTable seg_strided(int K, int Cp) {
  SegTable s;
  memset(&s, 0, sizeof(s));
  const int padL = (K - 2) / 2;
  s.nphase = 1;
  s.nseg[0] = K;
  for (int k = 0; k < K; ++k) {
    const int d = k - padL, j = floor_div2(d), p = d - 2 * j;
    s.shift[0][k] = (short)j;
    s.acol[0][k] = p * Cp;
    s.wk[0][k] = k * Cp;
  }
  return s;
}
// transposed form: out[b, 2q+r] = sum_{k : (r+padL-k) even} A[b, q + (r+padL-k)/2] * W[k]
static SegTable seg_transposed(int K, int Cp) {
  SegTable s;
  memset(&s, 0, sizeof(s));
  const int padL = (K - 2) / 2;
  s.nphase = 2;
  for (int r = 0; r < 2; ++r) {
    int n = 0;
    for (int k = 0; k < K; ++k) {
      const int e = r + padL - k;
      if (e & 1) continue;
      s.shift[r][n] = (short)(e / 2);
      s.acol[r][n] = 0;
      s.wk[r][n] = k * Cp;
      ++n;
    }
    s.nseg[r] = n;
  }
  return s;
}
static SegTable seg_dense() {
  SegTable s;
  memset(&s, 0, sizeof(s));
  s.nphase = 1;
  s.nseg[0] = 1;
  return s;
}

// ------------------------------------------------------------------------------------------ launchers
#define DISPATCH_T(c, ...)                    \
  do {                                        \
    if ((c)->bf) { using T = bf16; __VA_ARGS__; } \
    else { using T = float; __VA_ARGS__; }    \
  } while (0)

static int post_launch(cg_ctx* c, const char* what) {
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err("launch %s failed: %s", what, cudaGetErrorString(e));
  if (c->glue_open) {   // live timing of a memory-bound kernel (cg_profile): named after the launch that closes it
    c->glue_open = false;
    snprintf(c->prof.back().desc, sizeof(c->prof.back().desc), "%s", what);
    if (cudaEventRecord(c->prof.back().e1, c->stream) != cudaSuccess) return set_err("cudaEventRecord failed");
  }
  return 0;
}

static int prof_begin(cg_ctx* c, int cls, double flops, const char* desc = "") {
  if (!c->profiling) return 0;
  cg_ctx::ProfRec r;
  r.cls = cls; r.flops = cls == 3 ? 0.0 : flops; r.bytes = cls == 3 ? flops : 0.0;
  snprintf(r.desc, sizeof(r.desc), "%s", desc);
  CU(cudaEventCreate(&r.e0));
  CU(cudaEventCreate(&r.e1));
  CU(cudaEventRecord(r.e0, c->stream));
  c->prof.push_back(r);
  return 0;
}
static int prof_end(cg_ctx* c) {
  if (!c->profiling) return 0;
  CU(cudaEventRecord(c->prof.back().e1, c->stream));
  return 0;
}

// Launch of a memory-bound kernel, optionally (CG_GLUE_PDL=1) with the programmatic-dependent-launch attribute (its first
// statement is pdl_enter()): the kernel is then scheduled while its predecessor -- usually a persistent tensor-core kernel
// that triggered early -- drains. Measured same-box: 11.75 / 11.91 ms per step with it, 11.79 / 11.87 without, i.e. nothing
// (the stream has no launch gaps to recover at batch 128), so it stays off by default; the tensor-core kernels keep theirs.
template <typename... KArgs, typename... Args>
static inline void glue_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  static const bool pdl = getenv("CG_GLUE_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors surface through cudaGetLastError in post_launch
}

// opens a timing record for the NEXT launch of a memory-bound kernel; `bytes` = its algorithmic HBM traffic
static int glue(cg_ctx* c, double bytes) {
  if (!c->profiling) return 0;
  CK(prof_begin(c, 3, bytes, ""));
  c->glue_open = true;
  return 0;
}

static int launch_rsgemm_raw(cg_ctx* c, const RsParams& p) {
  if (c->use_tc && tc_rsgemm_supported(p)) {
    CK(tc_rsgemm_launch(&c->tc, p, c->stream));
    c->tc_launches++;
    return post_launch(c, "rsgemm_tc");
  }
  const long long M = (long long)p.B * p.Q;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)(p.N / 64), (unsigned)p.seg.nphase);
  DISPATCH_T(c, rsgemm_simt_kernel<T><<<grid, 256, 0, c->stream>>>(p));
  return post_launch(c, "rsgemm_simt");
}
static int launch_rsgemm(cg_ctx* c, const RsParams& p) {
  int nseg = 0;
  for (int i = 0; i < p.seg.nphase; ++i) nseg += p.seg.nseg[i];
  char d[96];
  snprintf(d, sizeof(d), "gemm  B=%d Q=%d N=%d Kc=%d taps=%d ph=%d epi=%d", p.B, p.Q, p.N, p.Kc, nseg, p.seg.nphase, p.epi);
  CK(prof_begin(c, 0, 2.0 * p.B * p.Q * (double)p.n_real * p.k_real * nseg * (p.flop_scale > 0.f ? p.flop_scale : 1.f), d));
  CK(launch_rsgemm_raw(c, p));
  return prof_end(c);
}

static int launch_wgrad_raw(cg_ctx* c, WgParams p);
static int launch_wgrad(cg_ctx* c, const WgParams& p) {
  char d[96];
  snprintf(d, sizeof(d), "wgrad B=%d Q=%d M=%d N=%d taps=%d", p.B, p.Q, p.Mp, p.Np, p.nseg);
  CK(prof_begin(c, 1, 2.0 * p.B * p.Q * (double)p.m_real * p.n_real * p.nseg, d));
  CK(launch_wgrad_raw(c, p));
  return prof_end(c);
}
static int launch_wgrad_raw(cg_ctx* c, WgParams p) {
  if (c->use_tc && tc_wgrad_supported(p)) {
    CK(tc_wgrad_launch(&c->tc, p, c->stream));
    c->tc_launches++;
    return post_launch(c, "wgrad_tc");
  }
  const long long R = (long long)p.B * p.Q;
  const int tiles = (p.Mp / 64) * (p.Np / 64) * p.nseg;
  int splits = (148 * 8 + tiles - 1) / tiles;
  long long max_splits = (R + 255) / 256;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  long long rps = (R + splits - 1) / splits;
  rps = (rps + 15) / 16 * 16;
  splits = (int)((R + rps - 1) / rps);
  p.rows_per_split = (int)rps;
  dim3 grid((unsigned)((p.Mp / 64) * (p.Np / 64)), (unsigned)p.nseg, (unsigned)splits);
  DISPATCH_T(c, wgrad_simt_kernel<T><<<grid, 256, 0, c->stream>>>(p));
  return post_launch(c, "wgrad_simt");
}

static void add_pack(PackOps& ops, const float* src, void* dst, int N, int n_real, int nseg, int Cp, int c_real,
                     long long sk, long long sn, long long sc, long long ld = 0, int k0 = 0, int kstep = 1, int slot0 = 0,
                     int slotstep = 1) {
  PackOp& o = ops.op[ops.n++];
  o.src = src; o.dst = dst; o.N = N; o.n_real = n_real; o.nseg = nseg; o.Cp = Cp; o.c_real = c_real;
  o.sk = sk; o.sn = sn; o.sc = sc; o.ld = ld ? ld : (long long)nseg * Cp;
  o.k0 = k0; o.kstep = kstep; o.slot0 = slot0; o.slotstep = slotstep;
}

static void* off(cg_ctx* c, void* base, long long elems) { return (char*)base + elems * c->esz; }

// refresh packed T copies of one model's GEMM weights from the fp32 master (one launch)
static int repack(cg_ctx* c, int which) {
  const int K = c->K;
  PackOps ops;
  ops.n = 0;
  if (which == CG_DISCRIMINATOR) {
    for (int l = 1; l <= NL; ++l) {
      const float* w = c->dis.w + c->dis.params[2 * (l - 1)].offset;   // (K, Cin, Cout)
      const int ci = c->dc[l - 1], co = c->dc[l], cip = c->dcp[l - 1], cop = c->dcp[l];
      add_pack(ops, w, c->Wf_d[l], cop, co, K, cip, ci, (long long)ci * co, 1, co);
      add_pack(ops, w, c->Wb_d[l], cip, ci, K, cop, co, (long long)ci * co, co, 1);
      if (c->Wf2_d[l]) {   // row-pair form: rows [g*64, g*64+64) hold the kernel delayed by 2g taps (the other taps stay zero)
        const long long ld2 = (long long)(K + 2) * cip;
        for (int g = 0; g < 2; ++g)
          add_pack(ops, w, off(c, c->Wf2_d[l], (long long)g * 64 * ld2 + 2LL * g * cip), cop, co, K, cip, ci, (long long)ci * co, 1,
                   co, ld2);
      }
      if (c->Wb2_d[l]) {   // merged phases: rows [g*64, g*64+64) = the taps of output phase g, each in its input window's slot
        const int padL = (K - 2) / 2;
        const long long ld2 = (long long)c->mp_K2 * cop;
        for (int g = 0; g < 2; ++g) {
          const int k0 = (padL & 1) == g ? 0 : 1;                    // first tap with (padL - k) & 1 == g
          const int d0 = padL - k0, slot0 = ((d0 + g) >> 1) - c->mp_emin;
          add_pack(ops, w, off(c, c->Wb2_d[l], (long long)g * 64 * ld2), cip, ci, (K - k0 + 1) / 2, cop, co, (long long)ci * co, co, 1,
                   ld2, k0, 2, slot0, -1);
        }
      }
    }
  } else {
    for (int i = 1; i <= NL; ++i) {
      const float* w = c->gen.w + c->gen.params[c->g_k[i]].offset;     // (K, 1, Cout, Cin)
      const int ci = c->gc[i - 1], co = c->gc[i], cip = c->gcp[i - 1], cop = c->gcp[i];
      add_pack(ops, w, c->Wf_g[i], cop, co, K, cip, ci, (long long)ci * co, ci, 1);
      add_pack(ops, w, c->Wb_g[i], cip, ci, K, cop, co, (long long)ci * co, 1, ci);
    }
    const float* w1 = c->gen.w + c->gen.params[c->g_d1k].offset;       // (C_in, C_out)
    const int C = c->C, Cp = c->gcp[NL];
    add_pack(ops, w1, c->Wf_d1, Cp, C, 1, Cp, C, 0, 1, C);             // [n=out][c=in]
    add_pack(ops, w1, c->Wb_d1, Cp, C, 1, Cp, C, 0, C, 1);             // [n=in][c=out]
  }
  if (ops.n > 24) return set_err("repack: too many pack ops");
  dim3 grid(16, c->K, ops.n);
  CK(glue(c, (4.0 + 2.0 * c->esz) * (which == CG_GENERATOR ? c->gen.total : c->dis.total)));
  DISPATCH_T(c, pack_weights_kernel<T><<<grid, 256, 0, c->stream>>>(ops));
  return post_launch(c, "pack_weights");
}

// Strided critic conv l in row-pair form (RsParams.row_pairs): 64 (padded) output channels, tensor-core path, and at
// least 128 GEMM rows (pairs of output time steps) per sample in whole 128-row blocks.
static bool row_pairs_apply(const cg_ctx* c, int l) {
  if (!c->use_tc || getenv("CG_NO_ROW_PAIRS")) return false;
  return c->dcp[l] == 64 && c->dl[l] % 256 == 0 && c->K + 2 <= CG_MAX_SEG;
}

// Data gradient of critic conv l with merged output phases (RsParams.merged_phases): 64 (padded) input channels, tensor-
// core path, whole 128-row blocks per sample.
static bool merged_phases_apply(const cg_ctx* c, int l) {
  if (!c->use_tc || getenv("CG_NO_MERGED_PHASES") || l < 2) return false;
  return c->dcp[l - 1] == 64 && c->dl[l] >= 128 && c->dl[l] % 128 == 0 && c->mp_K2 <= 32;
}

// one-pass Adam + re-pack: GEMM kernels as 32 x 32 tiles per tap, everything between them as element-wise ranges
static void build_adam_plans(cg_ctx* c) {
  for (int which = 0; which < 2; ++which) {
    AdamPlan& pl = c->adam_plan[which];
    memset(&pl, 0, sizeof(pl));
    Model& M = which == CG_GENERATOR ? c->gen : c->dis;
    std::vector<bool> is_gemm(M.params.size(), false);
    auto add_tensor = [&](int idx, int K, int A, int B, int Ap, int Bp, void* direct, void* trans, void* trans2 = nullptr,
                          void* direct2 = nullptr) {
      AdamTensor& t = pl.t[pl.nt++];
      t.off = M.params[idx].offset; t.K = K; t.A = A; t.B = B; t.Ap = Ap; t.Bp = Bp; t.direct = direct; t.trans = trans;
      t.trans2 = trans2; t.direct2 = direct2; t.d2_emin = c->mp_emin; t.d2_K2 = c->mp_K2;
      t.tiles_a = (A + 31) / 32; t.tiles_b = (B + 31) / 32;
      t.item0 = pl.tile_items;
      pl.tile_items += (long long)K * t.tiles_a * t.tiles_b;
      is_gemm[idx] = true;
    };
    if (which == CG_DISCRIMINATOR) {
      for (int l = 1; l <= NL; ++l)   // (K, Cin, Cout): a = ci, b = co; Wb_d = [ci][k*Coutp + co], Wf_d = [co][k*Cinp + ci]
        add_tensor(2 * (l - 1), c->K, c->dc[l - 1], c->dc[l], c->dcp[l - 1], c->dcp[l], c->Wb_d[l], c->Wf_d[l], c->Wf2_d[l],
                   c->Wb2_d[l]);
    } else {
      for (int i = 1; i <= NL; ++i)   // (K, 1, Cout, Cin): a = co, b = ci; Wf_g = [co][k*Cinp + ci], Wb_g = [ci][k*Coutp + co]
        add_tensor(c->g_k[i], c->K, c->gc[i], c->gc[i - 1], c->gcp[i], c->gcp[i - 1], c->Wf_g[i], c->Wb_g[i]);
      // output dense (C_in, C_out): a = in, b = out; Wb_d1 = [in][out], Wf_d1 = [out][in]
      add_tensor(c->g_d1k, 1, c->C, c->C, c->gcp[NL], c->gcp[NL], c->Wb_d1, c->Wf_d1);
    }
    long long items = pl.tile_items;
    for (size_t i = 0; i < M.params.size();) {
      if (is_gemm[i]) { ++i; continue; }
      size_t j = i;
      long long len = 0;
      while (j < M.params.size() && !is_gemm[j]) { len += M.params[j].size; ++j; }
      AdamRange& r = pl.r[pl.nr++];
      r.off = M.params[i].offset; r.len = len; r.item0 = items - pl.tile_items;
      items += (len + 1023) / 1024;
      i = j;
    }
    pl.items = items;
  }
}

// ------------------------------------------------------------------------------------------ create
extern "C" int cg_version(void) { return CG_VERSION; }
extern "C" const char* cg_last_error(void) { return g_err.c_str(); }

extern "C" void cg_destroy(cg_ctx* c) {
  if (!c) return;
  cudaDeviceSynchronize();
  tc_destroy(&c->tc);
  for (void* p : c->allocs) cudaFree(p);
  for (int w = 0; w < 2; ++w)
    for (int b = 0; b < cg_ctx::NB; ++b)
      if (c->bucket_evt[w][b]) cudaEventDestroy(c->bucket_evt[w][b]);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->side) cudaStreamDestroy(c->side);
  delete c;
}

extern "C" int cg_create(const cg_config* cfg, cg_ctx** out) {
  if (!cfg || !out) return set_err("cg_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_err("cg_create: no CUDA device (this library has no CPU fallback)");
  if (cfg->strides != 2) return set_err("cg_create: only strides == 2 is implemented (got %d)", cfg->strides);
  if (cfg->kernel_size < 2 || cfg->kernel_size > CG_MAX_SEG)
    return set_err("cg_create: kernel_size must be in [2, %d]", CG_MAX_SEG);
  if (cfg->seq_len % 32 != 0 || cfg->seq_len < 32)
    return set_err("Conv1D: w %g is not an integer.", cfg->seq_len / 32.0);   // calciumgan.py:17-18
  if (cfg->max_batch < 1 || cfg->channels < 1 || cfg->noise_dim < 1 || cfg->num_units < 1)
    return set_err("cg_create: bad sizes");
  if (cfg->phase_m < 0 || cfg->phase_m > cfg->seq_len / 32 * 2 - 1)
    return set_err("cg_create: phase shuffle m=%d too large for the shortest layer", cfg->phase_m);

  cg_ctx* c = new cg_ctx();
  c->cfg = *cfg;
  if (c->cfg.world_size < 1) c->cfg.world_size = 1;
  if (c->cfg.n_critic < 1) c->cfg.n_critic = 1;
  c->dbg_flags = cfg->debug_flags;
  if (getenv("CG_NO_PS_FUSE")) c->dbg_flags |= CG_DEBUG_NO_PS_FUSE;
  if (getenv("CG_NO_PS_BWD_FUSE")) c->dbg_flags |= CG_DEBUG_NO_PS_BWD_FUSE;
  if (getenv("CG_NO_GHEAD")) c->dbg_flags |= CG_DEBUG_NO_GHEAD;
  if (getenv("CG_NO_ADAM_FUSE")) c->dbg_flags |= CG_DEBUG_NO_ADAM_FUSE;
  c->bf = cfg->precision == CG_BF16;
  c->esz = c->bf ? 2 : 4;
  c->use_tc = c->bf && !cfg->force_simt;
  c->L = cfg->seq_len; c->C = cfg->channels; c->nd = cfg->noise_dim; c->nu = cfg->num_units;
  c->K = cfg->kernel_size; c->w0 = c->L / 32; c->Bmax = cfg->max_batch;
  const int nu = c->nu;
  const int gcs[NL + 1] = {c->nd, nu * 5, nu * 4, nu * 3, nu * 2, c->C};
  const int dcs[NL + 1] = {c->C, nu, nu * 2, nu * 3, nu * 4, nu * 5};
  for (int i = 0; i <= NL; ++i) {
    c->gc[i] = gcs[i]; c->gcp[i] = round_up(gcs[i], CG_CPAD); c->gl[i] = c->w0 << i;
    c->dc[i] = dcs[i]; c->dcp[i] = round_up(dcs[i], CG_CPAD); c->dl[i] = c->L >> i;
  }
  for (int i = 0; i <= NL; ++i)
    if (c->gcp[i] > 32 * 4 * (16 / c->esz)) {
      const int bad = c->gcp[i];
      delete c;
      return set_err("cg_create: %d generator channels exceed the layer-norm kernel's row width", bad);
    }
  // ---- parameter tables in Keras get_weights() order
  Model& G = c->gen;
  Model& D = c->dis;
  G.add({c->nd, (int64_t)c->w0 * c->nd}); G.add({(int64_t)c->w0 * c->nd});
  for (int i = 1; i <= NL; ++i) {
    c->g_k[i] = (int)G.params.size(); G.add({c->K, 1, c->gc[i], c->gc[i - 1]});
    c->g_b[i] = (int)G.params.size(); G.add({c->gc[i]});
    if (cfg->layer_norm) {
      c->g_gam[i] = (int)G.params.size(); G.add({c->gc[i]});
      c->g_bet[i] = (int)G.params.size(); G.add({c->gc[i]});
    }
  }
  c->g_d1k = (int)G.params.size(); G.add({c->C, c->C});
  c->g_d1b = (int)G.params.size(); G.add({c->C});
  for (int l = 1; l <= NL; ++l) { D.add({c->K, c->dc[l - 1], c->dc[l]}); D.add({c->dc[l]}); }
  D.add({(int64_t)c->dl[NL] * c->dc[NL], 1}); D.add({1});

 // bucket ranges: generator {[convT5 .. end), [convT3 .. convT5), [0 .. convT3)}; critic {[conv5 .. end), [conv4 .. conv5), [0 .. conv4)}
  c->bucket_off[CG_GENERATOR][0] = G.params[c->g_k[5]].offset; c->bucket_off[CG_GENERATOR][1] = G.params[c->g_k[3]].offset;
  c->bucket_off[CG_GENERATOR][2] = 0; c->bucket_off[CG_GENERATOR][3] = G.total;
  c->bucket_off[CG_DISCRIMINATOR][0] = D.params[8].offset; c->bucket_off[CG_DISCRIMINATOR][1] = D.params[6].offset;
  c->bucket_off[CG_DISCRIMINATOR][2] = 0; c->bucket_off[CG_DISCRIMINATOR][3] = D.total;
  for (int w = 0; w < 2; ++w)
    for (int b = 0; b < cg_ctx::NB; ++b)
      if (cudaEventCreateWithFlags(&c->bucket_evt[w][b], cudaEventDisableTiming) != cudaSuccess) {
        delete c;
        return set_err("cudaEventCreate failed");
      }
#define DA_(ptr, bytes)                                          \
  do {                                                           \
    if (dalloc(c, (void**)&(ptr), (size_t)(bytes))) { cg_destroy(c); return 1; } \
  } while (0)
  for (Model* m : {&G, &D}) {
    DA_(m->w, m->total * 4); DA_(m->m, m->total * 4); DA_(m->v, m->total * 4); DA_(m->g, m->total * 4);
    DA_(m->opt, sizeof(OptState));
    m->g_own = m->g;
  }
  const size_t es = c->esz;
  const size_t Bt = 3 * (size_t)c->Bmax, Bm = c->Bmax;
  for (int l = 1; l <= NL; ++l) {
    DA_(c->Wf_d[l], (size_t)c->dcp[l] * c->K * c->dcp[l - 1] * es);
    DA_(c->Wb_d[l], (size_t)c->dcp[l - 1] * c->K * c->dcp[l] * es);
    DA_(c->Wf_g[l], (size_t)c->gcp[l] * c->K * c->gcp[l - 1] * es);
    DA_(c->Wb_g[l], (size_t)c->gcp[l - 1] * c->K * c->gcp[l] * es);
  }
  {
    const int padL = (c->K - 2) / 2, dlo = padL - (c->K - 1);
    c->mp_emin = (dlo + (dlo & 1)) >> 1;                              // K = 24: windows e = -6 .. 6
    c->mp_K2 = ((padL + (padL & 1)) >> 1) - c->mp_emin + 1;
  }
  for (int l = 1; l <= NL; ++l) {
    c->Wf2_d[l] = c->Wb2_d[l] = nullptr;
    if (row_pairs_apply(c, l)) DA_(c->Wf2_d[l], (size_t)128 * (c->K + 2) * c->dcp[l - 1] * es);
    if (merged_phases_apply(c, l)) DA_(c->Wb2_d[l], (size_t)128 * c->mp_K2 * c->dcp[l] * es);
  }
  DA_(c->Wf_d1, (size_t)c->gcp[NL] * c->gcp[NL] * es);
  DA_(c->Wb_d1, (size_t)c->gcp[NL] * c->gcp[NL] * es);
  DA_(c->X[0], Bt * c->dl[0] * c->dcp[0] * es);
  DA_(c->DX[0], Bm * c->dl[0] * c->dcp[0] * es);
  c->H[0] = nullptr; c->DA[0] = nullptr;
  for (int l = 1; l <= NL; ++l) {
    const size_t n = Bt * c->dl[l] * c->dcp[l] * es;
    DA_(c->H[l], n); DA_(c->DA[l], n);
    if (l < NL) { DA_(c->X[l], n); DA_(c->DX[l], n); } else { c->X[l] = c->H[l]; c->DX[l] = nullptr; }
  }
  DA_(c->V5, Bm * c->dl[NL] * c->dcp[NL] * es);
  DA_(c->scores, Bt * 4); DA_(c->coef, Bt * 4); DA_(c->sumsq, Bm * 8); DA_(c->ucoef, Bm * 4); DA_(c->norms, Bm * 4);
  DA_(c->Z, Bm * c->nd * 4);
  for (int i = 0; i <= NL; ++i) {
    const size_t n = Bm * c->gl[i] * c->gcp[i] * es;
    DA_(c->HG[i], n); DA_(c->DHG[i], n);
    if (i >= 1) {
      DA_(c->AG[i], n); DA_(c->DAG[i], n);
      DA_(c->MU[i], Bm * c->gl[i] * 4); DA_(c->RSTD[i], Bm * c->gl[i] * 4);
    } else { c->AG[i] = c->DAG[i] = nullptr; c->MU[i] = c->RSTD[i] = nullptr; }
  }
  DA_(c->DO, Bm * c->L * c->gcp[NL] * es);
  DA_(c->FAKE32, Bm * c->L * c->C * 4);
  DA_(c->alpha_buf, (size_t)c->cfg.n_critic * Bm * 4);
  DA_(c->noise_buf, (size_t)(c->cfg.n_critic + 1) * Bm * c->nd * 4);
  DA_(c->d_scal, (size_t)(c->cfg.n_critic + 1) * CG_NUM_SCALARS * 4);
#undef DA_
  if (cudaMallocHost((void**)&c->h_scal, (size_t)(c->cfg.n_critic + 1) * CG_NUM_SCALARS * 4) != cudaSuccess) {
    cg_destroy(c);
    return set_err("cudaMallocHost failed");
  }
  build_adam_plans(c);
  if (c->use_tc && tc_init(&c->tc)) { cg_destroy(c); return 1; }
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) { cg_destroy(c); return set_err("arena init failed"); }
  *out = c;
  return 0;
}

extern "C" int cg_set_stream(cg_ctx* c, void* s) { c->stream = (cudaStream_t)s; return 0; }
extern "C" int cg_synchronize(cg_ctx* c) { CU(cudaStreamSynchronize(c->stream)); return 0; }
extern "C" int64_t cg_launch_count(cg_ctx* c) { return c->launches; }
extern "C" int64_t cg_tc_launch_count(cg_ctx* c) { return c->tc_launches; }
extern "C" int64_t cg_device_bytes(cg_ctx* c) { return c->dev_bytes; }
extern "C" void* cg_fake_ptr(cg_ctx* c) { return c->FAKE32; }
extern "C" void* cg_scores_ptr(cg_ctx* c) { return c->scores; }
extern "C" void* cg_scalars_ptr(cg_ctx* c) { return c->d_scal; }

static Model* model_of(cg_ctx* c, int which) { return which == CG_GENERATOR ? &c->gen : &c->dis; }
extern "C" int64_t cg_num_params(cg_ctx* c, int which) { return model_of(c, which)->total; }
extern "C" int cg_num_tensors(cg_ctx* c, int which) { return (int)model_of(c, which)->params.size(); }
extern "C" int cg_tensor_info(cg_ctx* c, int which, int idx, int64_t shape[4], int* ndim, int64_t* offset) {
  Model* m = model_of(c, which);
  if (idx < 0 || idx >= (int)m->params.size()) return set_err("cg_tensor_info: index %d out of range", idx);
  for (int i = 0; i < 4; ++i) shape[i] = m->params[idx].shape[i];
  *ndim = m->params[idx].ndim;
  *offset = m->params[idx].offset;
  return 0;
}
extern "C" int cg_set_weights(cg_ctx* c, int which, const float* host) {
  Model* m = model_of(c, which);
  CU(cudaMemcpyAsync(m->w, host, m->total * 4, cudaMemcpyHostToDevice, c->stream));
  CK(repack(c, which));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int cg_get_weights(cg_ctx* c, int which, float* host) {
  Model* m = model_of(c, which);
  CU(cudaMemcpyAsync(host, m->w, m->total * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int cg_get_grads(cg_ctx* c, int which, float* host) {
  Model* m = model_of(c, which);
  CU(cudaMemcpyAsync(host, m->g, m->total * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" void* cg_grad_ptr(cg_ctx* c, int which) { return model_of(c, which)->g; }
extern "C" int cg_num_buckets(cg_ctx*, int) { return cg_ctx::NB; }
extern "C" int cg_bucket_info(cg_ctx* c, int which, int bucket, int64_t* offset, int64_t* count) {
  if (which < 0 || which > 1 || bucket < 0 || bucket >= cg_ctx::NB) return set_err("cg_bucket_info: bad arguments");
  // bucket b spans [off[b], end_b) where end_0 = total and end_b = off[b-1]
  const int64_t lo = c->bucket_off[which][bucket];
  const int64_t hi = bucket == 0 ? c->bucket_off[which][cg_ctx::NB] : c->bucket_off[which][bucket - 1];
  *offset = lo;
  *count = hi - lo;
  return 0;
}
extern "C" int cg_stream_wait_bucket(cg_ctx* c, int which, int bucket, void* stream) {
  if (which < 0 || which > 1 || bucket < 0 || bucket >= cg_ctx::NB) return set_err("cg_stream_wait_bucket: bad arguments");
  CU(cudaStreamWaitEvent((cudaStream_t)stream, c->bucket_evt[which][bucket], 0));
  return 0;
}
extern "C" int cg_get_opt_state(cg_ctx* c, int which, float* hm, float* hv, int64_t* step) {
  Model* m = model_of(c, which);
  OptState st;
  if (hm) CU(cudaMemcpyAsync(hm, m->m, m->total * 4, cudaMemcpyDeviceToHost, c->stream));
  if (hv) CU(cudaMemcpyAsync(hv, m->v, m->total * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&st, m->opt, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (step) *step = st.steps;
  return 0;
}
extern "C" int cg_set_opt_state(cg_ctx* c, int which, const float* hm, const float* hv, int64_t step) {
  Model* m = model_of(c, which);
  if (hm) CU(cudaMemcpyAsync(m->m, hm, m->total * 4, cudaMemcpyHostToDevice, c->stream));
  if (hv) CU(cudaMemcpyAsync(m->v, hv, m->total * 4, cudaMemcpyHostToDevice, c->stream));
  OptState st;
  memset(&st, 0, sizeof(st));
  CU(cudaMemcpyAsync(&st, m->opt, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  st.steps = step;
  CU(cudaMemcpyAsync(m->opt, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int64_t cg_skipped_updates(cg_ctx* c, int which) {
  OptState st;
  memset(&st, 0, sizeof(st));
  if (cudaMemcpyAsync(&st, model_of(c, which)->opt, sizeof(st), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
      cudaStreamSynchronize(c->stream) != cudaSuccess)
    return -1;
  return st.skipped;
}
// Data parallel over NVLink peer memory: gradients are accumulated straight into a caller-provided buffer that every
// peer has mapped (symmetric memory); NULL restores the library's own buffer.
extern "C" int cg_set_grad_buffer(cg_ctx* c, int which, float* dev) {
  Model* m = model_of(c, which);
  if (dev && (reinterpret_cast<uintptr_t>(dev) & 15)) return set_err("cg_set_grad_buffer: the buffer must be 16-byte aligned");
  CU(cudaStreamSynchronize(c->stream));
  m->g = dev ? dev : m->g_own;
  return 0;
}
static int peer_ptrs(const void* const* host, int world, PeerPtrs& pp) {
  if (!host || (world != 2 && world != 4 && world != 8)) return set_err("peer gradient exchange: world must be 2, 4 or 8");
  for (int r = 0; r < 8; ++r) {
    pp.p[r] = (const float*)(r < world ? host[r] : host[0]);
    if (!pp.p[r] || (reinterpret_cast<uintptr_t>(pp.p[r]) & 15)) return set_err("peer gradient exchange: peer buffers must be 16-byte aligned");
  }
  return 0;
}
static int reduced_buffer(cg_ctx* c, Model* m) {
  if (m->gr) return 0;
  if (dalloc(c, (void**)&m->gr, (size_t)(m->total + 4) * 4)) return 1;
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
static long long peer_slice(const Model* m, int world) { return ((m->total + world - 1) / world + 3) / 4 * 4; }
static int launch_peer_sum(cg_ctx* c, const PeerPtrs& pp, int world, float* out, long long first, long long n, cudaStream_t st) {
  static const int ctas = getenv("CG_PEER_CTAS") ? atoi(getenv("CG_PEER_CTAS")) : 64;
  long long g = (n / 4 + 63) / 64;
  const int grid = (int)(g < 1 ? 1 : (g > ctas ? ctas : g));
  if (world == 2) peer_sum_kernel<2, 4><<<grid, 64, 0, st>>>(pp, out, first, n);
  else if (world == 4) peer_sum_kernel<4, 2><<<grid, 64, 0, st>>>(pp, out, first, n);
  else peer_sum_kernel<8, 1><<<grid, 64, 0, st>>>(pp, out, first, n);
  return post_launch(c, "peer_sum");
}
// one-shot: reduced (read by cg_apply_update_reduced) = sum over ranks r = 0 .. world-1 of peer_ptrs[r][0 .. num_params),
// launched on `cuda_stream`. The caller orders it between two cross-rank barriers (all gradients written / all peers
// have finished reading) on the same stream. Pulls (world - 1) x the gradient size: the right form for 2 ranks.
extern "C" int cg_reduce_peer_grads(cg_ctx* c, int which, const void* const* peer_ptrs_host, int world, void* cuda_stream) {
  Model* m = model_of(c, which);
  PeerPtrs pp;
  CK(peer_ptrs(peer_ptrs_host, world, pp));
  CK(reduced_buffer(c, m));
  return launch_peer_sum(c, pp, world, m->gr, 0, m->total, (cudaStream_t)cuda_stream);
}
// two-phase form for 4 / 8 ranks (2 x (world - 1) / world x the gradient size per rank instead of (world - 1) x):
// phase 1, reduce-scatter: this rank sums ITS slice of every peer's gradient buffer into its reduced buffer, which must
// be peer-mapped too (cg_set_reduced_buffer); phase 2, all-gather: it copies the other slices from their owners'
// reduced buffers. Barriers: before phase 1 (gradients complete) and between the phases (slices complete; gradient
// buffers may be overwritten); the next exchange's first barrier also protects the reduced slices.
extern "C" int cg_peer_reduce_scatter(cg_ctx* c, int which, const void* const* grad_peer_ptrs_host, int world, int rank,
                                      void* cuda_stream) {
  Model* m = model_of(c, which);
  PeerPtrs pp;
  CK(peer_ptrs(grad_peer_ptrs_host, world, pp));
  if (rank < 0 || rank >= world || !m->gr) return set_err("cg_peer_reduce_scatter: bad rank or no reduced buffer (cg_set_reduced_buffer)");
  const long long slice = peer_slice(m, world), first = (long long)rank * slice;
  long long cnt = m->total - first;
  if (cnt > slice) cnt = slice;
  if (cnt <= 0) return 0;
  return launch_peer_sum(c, pp, world, m->gr, first, cnt, (cudaStream_t)cuda_stream);
}
extern "C" int cg_peer_all_gather(cg_ctx* c, int which, const void* const* reduced_peer_ptrs_host, int world, int rank,
                                  void* cuda_stream) {
  Model* m = model_of(c, which);
  PeerPtrs pp;
  CK(peer_ptrs(reduced_peer_ptrs_host, world, pp));
  if (rank < 0 || rank >= world || !m->gr) return set_err("cg_peer_all_gather: bad rank or no reduced buffer (cg_set_reduced_buffer)");
  const long long slice = peer_slice(m, world);
  static const int ctas = getenv("CG_PEER_CTAS") ? atoi(getenv("CG_PEER_CTAS")) : 64;
  long long g = (slice / 4 + 63) / 64;
  int gx = ctas / world;
  if (gx < 1) gx = 1;
  if (gx > g) gx = (int)g;
  peer_gather_kernel<<<dim3(gx, world), 64, 0, (cudaStream_t)cuda_stream>>>(pp, m->gr, slice, m->total, rank);
  return post_launch(c, "peer_gather");
}
// the buffer cg_apply_update_reduced reads: caller-provided (peer-mapped, >= num_params + 4 floats) or, with NULL, the
// library's own
extern "C" int cg_set_reduced_buffer(cg_ctx* c, int which, float* dev) {
  Model* m = model_of(c, which);
  if (dev && (reinterpret_cast<uintptr_t>(dev) & 15)) return set_err("cg_set_reduced_buffer: the buffer must be 16-byte aligned");
  CU(cudaStreamSynchronize(c->stream));
  m->gr = dev;
  if (!dev) CK(reduced_buffer(c, m));
  return 0;
}
extern "C" void* cg_reduced_grad_ptr(cg_ctx* c, int which) { return model_of(c, which)->gr; }

extern "C" int cg_set_grads(cg_ctx* c, int which, const float* host) {
  Model* m = model_of(c, which);
  if (!host) return set_err("cg_set_grads: null pointer");
  CU(cudaMemcpyAsync(m->g, host, m->total * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int cg_seed(cg_ctx* c, uint64_t seed) {
  c->seed = seed;
  c->noise_calls = c->alpha_calls = c->shift_draws = 0;
  return 0;
}

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Keras glorot-uniform / zeros / LN (1,0) init on the host, then upload
extern "C" int cg_init_weights(cg_ctx* c, uint64_t seed) {
  uint64_t st = splitmix64(seed);
  auto uni = [&]() { st = splitmix64(st); return (double)(st >> 11) * (1.0 / 9007199254740992.0); };
  for (int which = 0; which < 2; ++which) {
    Model* m = model_of(c, which);
    std::vector<float> h((size_t)m->total, 0.f);
    for (size_t i = 0; i < m->params.size(); ++i) {
      const ParamInfo& p = m->params[i];
      double fan_in = 0, fan_out = 0;
      if (p.ndim == 2) { fan_in = (double)p.shape[0]; fan_out = (double)p.shape[1]; }
      else if (p.ndim == 3) { fan_in = (double)p.shape[0] * p.shape[1]; fan_out = (double)p.shape[0] * p.shape[2]; }
      else if (p.ndim == 4) { fan_in = (double)p.shape[0] * p.shape[3]; fan_out = (double)p.shape[0] * p.shape[2]; }
      if (p.ndim >= 2) {
        const double lim = std::sqrt(6.0 / (fan_in + fan_out));
        for (int64_t j = 0; j < p.size; ++j) h[p.offset + j] = (float)((2.0 * uni() - 1.0) * lim);
      }
    }
    if (which == CG_GENERATOR && c->cfg.layer_norm)
      for (int i = 1; i <= NL; ++i) {
        const ParamInfo& p = m->params[c->g_gam[i]];
        for (int64_t j = 0; j < p.size; ++j) h[p.offset + j] = 1.f;
      }
    CK(cg_set_weights(c, which, h.data()));
    CU(cudaMemsetAsync(m->m, 0, m->total * 4, c->stream));
    CU(cudaMemsetAsync(m->v, 0, m->total * 4, c->stream));
    CU(cudaMemsetAsync(m->opt, 0, sizeof(OptState), c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------ generator
static float* gparam(cg_ctx* c, int idx) { return c->gen.w + c->gen.params[idx].offset; }
static float* ggrad(cg_ctx* c, int idx) { return c->gen.g + c->gen.params[idx].offset; }
static float* dparam(cg_ctx* c, int idx) { return c->dis.w + c->dis.params[idx].offset; }
static float* dgrad(cg_ctx* c, int idx) { return c->dis.g + c->dis.params[idx].offset; }

// Conv1DTranspose i forward (models/utils.py:79-89): HG[i-1] -> AG[i] (+bias)
static RsParams convT_fwd_params(cg_ctx* c, int i, int B) {
  RsParams p;
  memset(&p, 0, sizeof(p));
  p.A = c->HG[i - 1]; p.a_bs = (long long)c->gl[i - 1] * c->gcp[i - 1]; p.a_rs = c->gcp[i - 1]; p.a_rows = c->gl[i - 1];
  p.W = c->Wf_g[i]; p.w_ld = c->K * c->gcp[i - 1];
  p.out = c->AG[i]; p.o_bs = (long long)c->gl[i] * c->gcp[i]; p.o_rs = 2 * c->gcp[i]; p.o_phase_col = c->gcp[i];
  p.bias = gparam(c, c->g_b[i]);
  p.B = B; p.Q = c->gl[i - 1]; p.N = c->gcp[i]; p.n_real = c->gc[i]; p.Kc = c->gcp[i - 1]; p.k_real = c->gc[i - 1];
  p.epi = EPI_BIAS;
  p.seg = seg_transposed(c->K, c->gcp[i - 1]);
  return p;
}
// data gradient of Conv1DTranspose i: dHG[i-1][b,q,ci] = sum_k DAG[i][b, 2q+k-padL, co] * Wt[k][co][ci]
static RsParams convT_bwd_params(cg_ctx* c, int i, int B) {
  RsParams p;
  memset(&p, 0, sizeof(p));
  p.A = c->DAG[i]; p.a_bs = (long long)c->gl[i] * c->gcp[i]; p.a_rs = 2 * c->gcp[i]; p.a_rows = c->gl[i] / 2;
  p.W = c->Wb_g[i]; p.w_ld = c->K * c->gcp[i];
  p.out = c->DHG[i - 1]; p.o_bs = (long long)c->gl[i - 1] * c->gcp[i - 1]; p.o_rs = c->gcp[i - 1];
  p.B = B; p.Q = c->gl[i - 1]; p.N = c->gcp[i - 1]; p.n_real = c->gc[i - 1]; p.Kc = c->gcp[i]; p.k_real = c->gc[i]; p.epi = EPI_NONE;
  p.seg = seg_strided(c->K, c->gcp[i]);
  return p;
}
// dWt[k][co][ci] = sum_{b,q} DAG[i][b, 2q + k - padL, co] * HG[i-1][b, q, ci]
static WgParams convT_wgrad_params(cg_ctx* c, int i, int B) {
  WgParams w;
  memset(&w, 0, sizeof(w));
  const SegTable st = seg_strided(c->K, c->gcp[i]);
  w.S = c->DAG[i]; w.s_bs = (long long)c->gl[i] * c->gcp[i]; w.s_rs = 2 * c->gcp[i]; w.s_rows = c->gl[i] / 2;
  w.P = c->HG[i - 1]; w.p_bs = (long long)c->gl[i - 1] * c->gcp[i - 1]; w.p_rs = c->gcp[i - 1];
  w.dW = ggrad(c, c->g_k[i]); w.m_real = c->gc[i]; w.n_real = c->gc[i - 1];
  w.B = B; w.Q = c->gl[i - 1]; w.Mp = c->gcp[i]; w.Np = c->gcp[i - 1]; w.nseg = c->K;
  for (int k = 0; k < c->K; ++k) { w.shift[k] = st.shift[0][k]; w.scol[k] = st.acol[0][k]; }
  return w;
}

// calciumgan.py:22-103. noise (B, nd) fp32 device. Writes FAKE32 (B, L, C).
// Optional fusion on the tensor-core path: xhat_slot != null also writes x_hat = alpha * real + (1 - alpha) * fake
// (wgan_gp.py:38-41) from the same epilogue; want32 = false skips the fp32 copy (critic sub-steps of a train step).
static int g_forward(cg_ctx* c, const float* noise, int B, void* fake_slot = nullptr, bool for_backward = true,
                     void* xhat_slot = nullptr, const float* real = nullptr, const float* alpha = nullptr,
                     bool want32 = true, bool* xhat_done = nullptr) {
  CK(glue(c, (double)B * c->nd * 4 + (double)c->nd * c->w0 * c->nd * 4 + (double)B * c->w0 * c->gcp[0] * c->esz));
  DISPATCH_T(c, glue_launch(dense0_forward_kernel<T>, dim3(grid_for((long long)B * c->w0 * c->gcp[0])), dim3(256), 0, c->stream, 
                    noise, gparam(c, 0), gparam(c, 1), (T*)c->HG[0], B, c->nd, c->w0, c->gcp[0]));
  CK(post_launch(c, "dense0_fwd"));
  for (int i = 1; i <= NL; ++i) {
    RsParams cp = convT_fwd_params(c, i, B);
    if (c->cfg.layer_norm && c->use_tc && !c->tc.force_v1 && tc_rsgemm2_supported(cp) && tc_ln_fusable(cp)) {
      // conv-transpose + bias + layer-norm + LeakyReLU in one kernel (row statistics are thread-local in the epilogue)
      cp.epi = EPI_BIAS_LN_LRELU;
      cp.out = c->HG[i];
      cp.gamma = gparam(c, c->g_gam[i]);
      cp.beta = gparam(c, c->g_bet[i]);
      if (for_backward) { cp.aux = c->AG[i]; cp.mu = c->MU[i]; cp.rstd = c->RSTD[i]; }
      CK(launch_rsgemm(c, cp));
      continue;
    }
    CK(launch_rsgemm(c, cp));
    const long long rows = (long long)B * c->gl[i];
    if (c->cfg.layer_norm) {
      const int nvec = c->gcp[i] / (16 / c->esz);
      const int lpr = nvec > 16 ? 32 : (nvec > 8 ? 16 : 8);
CK(glue(c, 2.0 * rows * c->gcp[i] * c->esz + 8.0 * rows));
#define CG_LN(LPRV)                                                                                              \
  DISPATCH_T(c, glue_launch(ln_lrelu_forward_kernel<T, LPRV>, dim3(grid_for(rows * LPRV)), dim3(256), 0, c->stream,                     \
                    (const T*)c->AG[i], gparam(c, c->g_gam[i]), gparam(c, c->g_bet[i]), (T*)c->HG[i], c->MU[i], \
                    c->RSTD[i], rows, c->gc[i], c->gcp[i]))
      if (lpr == 32) CG_LN(32); else if (lpr == 16) CG_LN(16); else CG_LN(8);
#undef CG_LN
      CK(post_launch(c, "ln_fwd"));
    } else {
      CK(glue(c, 2.0 * rows * c->gcp[i] * c->esz));
      DISPATCH_T(c, lrelu_kernel<T><<<grid_for(rows * c->gcp[i]), 256, 0, c->stream>>>((const T*)c->AG[i],
                                                                                      (T*)c->HG[i], rows * c->gcp[i]));
      CK(post_launch(c, "lrelu"));
    }
  }
  const int Cp = c->gcp[NL];
  if (xhat_done) *xhat_done = false;
  if (c->use_tc && !c->tc.force_v1 && !(c->dbg_flags & CG_DEBUG_NO_GHEAD)) {   // dedicated HBM-bound head kernel (cg_kernels_head.cuh)
    GHeadArgs a;
    memset(&a, 0, sizeof(a));
    a.A = c->HG[NL]; a.W = c->Wf_d1; a.bias = gparam(c, c->g_d1b);
    a.fake16 = fake_slot; a.xhat16 = xhat_slot; a.real = real; a.alpha = alpha;
    a.out32 = (want32 || !fake_slot) ? c->FAKE32 : nullptr;
    a.B = B; a.L = c->L; a.C = c->C; a.Cp = Cp; a.sigmoid = c->cfg.normalize ? 1 : 0;
    if (Cp == c->dcp[0] && tc_ghead_supported(a)) {
      char d[96];
      snprintf(d, sizeof(d), "ghead B=%d L=%d C=%d f=%d x=%d o32=%d", B, c->L, c->C, a.fake16 != nullptr, a.xhat16 != nullptr,
               a.out32 != nullptr);
      CK(prof_begin(c, 2, 2.0 * B * c->L * (double)c->C * c->C, d));
      if (c->profiling)   // HBM-bound: activations in, [real in], fake [+ x_hat] out in the compute type, [fp32 copy out]
        c->prof.back().bytes = (double)B * c->L * (Cp * 2.0 + (a.xhat16 ? 4.0 * c->C : 0.0) + (a.fake16 ? Cp * 2.0 : 0.0) +
                                                  (a.xhat16 ? Cp * 2.0 : 0.0) + (a.out32 ? 4.0 * c->C : 0.0));
      CK(tc_ghead_launch(&c->tc, a, c->stream));
      c->tc_launches++;
      CK(post_launch(c, "ghead_tc"));
      CK(prof_end(c));
      if (xhat_done) *xhat_done = xhat_slot != nullptr;
      return 0;
    }
  }
  RsParams p;
  memset(&p, 0, sizeof(p));
  p.A = c->HG[NL]; p.a_bs = (long long)c->L * Cp; p.a_rs = Cp; p.a_rows = c->L;
  p.W = c->Wf_d1; p.w_ld = Cp;
  p.out = fake_slot;     // optional compute-type copy (critic input slot; Cp == dcp[0])
  p.out32 = c->FAKE32; p.o32_bs = (long long)c->L * c->C; p.o32_rs = c->C;
  p.o_bs = (long long)c->L * Cp; p.o_rs = Cp;
  p.bias = gparam(c, c->g_d1b);
  p.B = B; p.Q = c->L; p.N = Cp; p.n_real = c->C; p.Kc = Cp; p.k_real = c->C;
  p.epi = c->cfg.normalize ? EPI_BIAS_SIGMOID : EPI_BIAS;
  if (const char* e = getenv("CG_HEAD_DBG")) p.dbg = atoi(e);
  p.seg = seg_dense();
  CK(launch_rsgemm(c, p));
  return 0;
}

static int launch_colsum_ops(cg_ctx* c, const ColsumOps& ops) {
  for (int i = 0; i < ops.n; ++i)
    if (ops.op[i].Cp / (16 / c->esz) > 256) return set_err("colsum: more than 256 16-byte vectors per row");
  dim3 grid(148 * 2, ops.n);
  double cs_bytes = 0;
  for (int i = 0; i < ops.n; ++i) cs_bytes += (double)ops.op[i].rows * ops.op[i].Cp * c->esz;
  CK(glue(c, cs_bytes));
  DISPATCH_T(c, glue_launch(colsum_multi_kernel<T>, dim3(grid), dim3(256), 0, c->stream, ops));
  return post_launch(c, "colsum");
}
static int launch_colsum(cg_ctx* c, const void* X, float* out, long long rows, int Cp, int c_real) {
  if (Cp / (16 / c->esz) > 256) {
    int gx = (int)((rows + 3) / 4);
    if (gx > 148 * 4) gx = 148 * 4;
    dim3 grid(gx, (c_real + 63) / 64), block(64, 4);
    CK(glue(c, (double)rows * Cp * c->esz));
    DISPATCH_T(c, colsum_kernel<T><<<grid, block, 0, c->stream>>>((const T*)X, out, rows, Cp, c_real));
    return post_launch(c, "colsum");
  }
  ColsumOps ops;
  ops.n = 1;
  ops.op[0].X = X; ops.op[0].out = out; ops.op[0].rows = rows; ops.op[0].Cp = Cp; ops.op[0].c_real = c_real;
  return launch_colsum_ops(c, ops);
}

// backward of the generator given DX[0] = dLoss/dfake (B, L, dcp0); accumulates into gen.g
static int g_backward(cg_ctx* c, int B) {
  const int Cp = c->gcp[NL];
  const long long rowsL = (long long)B * c->L;
  CK(glue(c, (double)rowsL * (2.0 * Cp * c->esz + 4.0 * c->C)));
  DISPATCH_T(c, glue_launch(sigmoid_backward_kernel<T>, dim3(grid_for(rowsL * Cp / (16 / c->esz))), dim3(256), 0, c->stream, 
                    (const T*)c->DX[0], c->FAKE32, (T*)c->DO, rowsL, c->C, Cp, c->cfg.normalize));
  CK(post_launch(c, "sigmoid_bwd"));
  {  // output dense: dW1[c_in][c_out] = sum_rows HG5[row,c_in] * DO[row,c_out]
    WgParams w;
    memset(&w, 0, sizeof(w));
    w.S = c->HG[NL]; w.s_bs = (long long)c->L * Cp; w.s_rs = Cp; w.s_rows = c->L;
    w.P = c->DO; w.p_bs = (long long)c->L * Cp; w.p_rs = Cp;
    w.dW = ggrad(c, c->g_d1k); w.m_real = c->C; w.n_real = c->C;
    w.B = B; w.Q = c->L; w.Mp = Cp; w.Np = Cp; w.nseg = 1;
    CK(launch_wgrad(c, w));
    CK(launch_colsum(c, c->DO, ggrad(c, c->g_d1b), rowsL, Cp, c->C));
    RsParams p;
    memset(&p, 0, sizeof(p));
    p.A = c->DO; p.a_bs = (long long)c->L * Cp; p.a_rs = Cp; p.a_rows = c->L;
    p.W = c->Wb_d1; p.w_ld = Cp;
    p.out = c->DHG[NL]; p.o_bs = (long long)c->L * Cp; p.o_rs = Cp;
    p.B = B; p.Q = c->L; p.N = Cp; p.n_real = c->C; p.Kc = Cp; p.k_real = c->C; p.epi = EPI_NONE;
    p.seg = seg_dense();
    CK(launch_rsgemm(c, p));
  }
  for (int i = NL; i >= 1; --i) {
    const long long rows = (long long)B * c->gl[i];
    if (c->cfg.layer_norm) {
      const int nvec_b = c->gcp[i] / (16 / c->esz);
      const int lpr_b = nvec_b > 16 ? 32 : (nvec_b > 8 ? 16 : 8);
      const int maxv = (nvec_b + lpr_b - 1) / lpr_b;   // channel vectors per lane: sizes the kernel's register arrays
      if (maxv > 4) return set_err("ln_lrelu_backward: more than 4 channel vectors per lane (Cp %d)", c->gcp[i]);
      const int blocks = grid_for(rows * lpr_b, 256, 148 * (maxv <= 1 ? 6 : 3));
CK(glue(c, 4.0 * rows * c->gcp[i] * c->esz + 8.0 * rows));
#define CG_LNB(LPRV, MV)                                                                                              \
  DISPATCH_T(c, glue_launch(ln_lrelu_backward_kernel<T, LPRV, MV>, dim3(blocks), dim3(256), 2 * c->gcp[i] * sizeof(float), c->stream,     \
                    (const T*)c->DHG[i], (const T*)c->AG[i], (const T*)c->HG[i], c->MU[i], c->RSTD[i],               \
                    gparam(c, c->g_gam[i]), (T*)c->DAG[i], ggrad(c, c->g_gam[i]), ggrad(c, c->g_bet[i]), rows,       \
                    c->gc[i], c->gcp[i]))
      if (lpr_b == 8) CG_LNB(8, 1);
      else if (lpr_b == 16) CG_LNB(16, 1);
      else if (maxv == 1) CG_LNB(32, 1);
      else if (maxv == 2) CG_LNB(32, 2);
      else if (maxv == 3) CG_LNB(32, 3);
      else CG_LNB(32, 4);
#undef CG_LNB
      CK(post_launch(c, "ln_bwd"));
    } else {
      CK(glue(c, 3.0 * rows * c->gcp[i] * c->esz));
      DISPATCH_T(c, mask_mul_kernel<T><<<grid_for(rows * c->gcp[i]), 256, 0, c->stream>>>(
                        (const T*)c->DHG[i], (const T*)c->HG[i], (T*)c->DAG[i], rows * c->gcp[i]));
      CK(post_launch(c, "mask_mul"));
    }
    CK(launch_wgrad(c, convT_wgrad_params(c, i, B)));
    CK(launch_colsum(c, c->DAG[i], ggrad(c, c->g_b[i]), rows, c->gcp[i], c->gc[i]));
    if (i == NL) CU(cudaEventRecord(c->bucket_evt[CG_GENERATOR][0], c->stream));       // output dense + convT5 done
    if (i == NL - 2) CU(cudaEventRecord(c->bucket_evt[CG_GENERATOR][1], c->stream));   // convT4, convT3 done
    CK(launch_rsgemm(c, convT_bwd_params(c, i, B)));
  }
  const int tot = (c->nd + 1) * c->w0 * c->nd;
  CK(glue(c, 2.0 * B * c->w0 * c->gcp[0] * c->esz + (double)B * c->nd * 4));
  DISPATCH_T(c, glue_launch(dense0_backward_kernel<T>, dim3(grid_for(tot)), dim3(256), 0, c->stream, 
                    c->Z, (const T*)c->DHG[0], (const T*)c->HG[0], ggrad(c, 0), ggrad(c, 1), B, c->nd, c->w0,
                    c->gcp[0]));
  CK(post_launch(c, "dense0_bwd"));
  CU(cudaEventRecord(c->bucket_evt[CG_GENERATOR][2], c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------ critic
static GroupShifts group_shifts(const int32_t* sh, int groups, int layer /*1..4*/) {
  GroupShifts g;
  for (int i = 0; i < 4; ++i) g.s[i] = i < groups ? sh[i * 4 + (layer - 1)] : 0;
  return g;
}

static RsParams conv_fwd_params(cg_ctx* c, int l, const void* A, void* out, int Bt, int epi, const void* mask) {
  RsParams p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.a_bs = (long long)c->dl[l - 1] * c->dcp[l - 1]; p.a_rs = 2 * c->dcp[l - 1]; p.a_rows = c->dl[l - 1] / 2;
  p.W = c->Wf_d[l]; p.w_ld = c->K * c->dcp[l - 1];
  p.out = out; p.o_bs = (long long)c->dl[l] * c->dcp[l]; p.o_rs = c->dcp[l];
  p.bias = dparam(c, 2 * (l - 1) + 1);
  p.mask = mask;
  p.B = Bt; p.Q = c->dl[l]; p.N = c->dcp[l]; p.n_real = c->dc[l]; p.Kc = c->dcp[l - 1]; p.k_real = c->dc[l - 1]; p.epi = epi;
  p.seg = seg_strided(c->K, c->dcp[l - 1]);
  if (c->Wf2_d[l] && !c->tc.force_v1) {
    // row-pair form: GEMM row i = output time steps (2i, 2i+1); input through the (B, L/4, 4*Cp) view; tap k' = k + 2g
    // of the widened kernel reads input time 4i + k' - padL for output 2i + g
    const int K2 = c->K + 2, Cp = c->dcp[l - 1], padL = (c->K - 2) / 2;
    p.a_rs = 4 * Cp; p.a_rows = c->dl[l - 1] / 4;
    p.W = c->Wf2_d[l]; p.w_ld = K2 * Cp;
    p.o_rs = 2 * c->dcp[l];
    p.Q = c->dl[l] / 2; p.N = 2 * c->dcp[l];
    p.row_pairs = 1;
    p.flop_scale = 2.0f * c->K / K2;
    memset(&p.seg, 0, sizeof(p.seg));
    p.seg.nphase = 1;
    p.seg.nseg[0] = K2;
    for (int k = 0; k < K2; ++k) {
      const int d = k - padL, j = d >= 0 ? d / 4 : -((-d + 3) / 4), r = d - 4 * j;
      p.seg.shift[0][k] = (short)j;
      p.seg.acol[0][k] = r * Cp;
      p.seg.wk[0][k] = k * Cp;
    }
  }
  return p;
}

// calciumgan.py:141-192 on X[0][0:Bt]; groups of B samples share PhaseShuffle shifts sh[g*4 + layer-1]
static bool ps_fusable(cg_ctx* c, const RsParams& p) {
  return !(c->dbg_flags & CG_DEBUG_NO_PS_FUSE) && c->use_tc && !c->tc.force_v1 && tc_rsgemm2_supported(p);
}
static void set_ps(RsParams& p, void* X, int w, int group_b, const int32_t* sh, int groups, int layer) {
  p.ps_out = X; p.ps_w = w; p.ps_group_b = group_b;
  for (int i = 0; i < 4; ++i) p.ps_shift[i] = i < groups ? sh[i * 4 + (layer - 1)] : 0;
}

static int d_forward(cg_ctx* c, int Bt, int B, int groups, const int32_t* sh) {
  for (int l = 1; l <= NL; ++l) {
    RsParams p = conv_fwd_params(c, l, c->X[l - 1], c->H[l], Bt, EPI_BIAS_LRELU, nullptr);
    if (l < NL && ps_fusable(c, p)) {   // conv + bias + LeakyReLU + PhaseShuffle in one kernel
      set_ps(p, c->X[l], c->dl[l], B, sh, groups, l);
      CK(launch_rsgemm(c, p));
      continue;
    }
    CK(launch_rsgemm(c, p));
    if (l < NL) {
      const long long tot = (long long)Bt * c->dl[l] * c->dcp[l] / (16 / c->esz);
      CK(glue(c, 2.0 * Bt * c->dl[l] * c->dcp[l] * c->esz));
      DISPATCH_T(c, ps_gather_kernel<T><<<grid_for(tot), 256, 0, c->stream>>>(
                        (const T*)c->H[l], (T*)c->X[l], Bt, B, c->dl[l], c->dcp[l], group_shifts(sh, groups, l)));
      CK(post_launch(c, "ps_gather"));
    }
  }
  CK(glue(c, (double)Bt * c->dl[NL] * c->dcp[NL] * c->esz + 4.0 * c->dl[NL] * c->dc[NL]));
  DISPATCH_T(c, glue_launch(head_forward_kernel<T>, dim3(Bt), dim3(256), 0, c->stream, (const T*)c->X[NL], dparam(c, 10), dparam(c, 11),
                                                                  c->scores, c->dl[NL], c->dc[NL], c->dcp[NL]));
  return post_launch(c, "head_fwd");
}

// data-gradient of conv layer l: DA[l] (rows dl[l]) -> out (rows dl[l-1]) for samples [b0, b0+nb)
static RsParams d_dgrad_params(cg_ctx* c, int l, int b0, int nb, void* out);
static bool d_dgrad_ps_fusable(cg_ctx* c, int l, int Bt) {
  if ((c->dbg_flags & CG_DEBUG_NO_PS_BWD_FUSE) || !c->use_tc || c->tc.force_v1 || !c->tc.use_pair || c->cfg.phase_m > 10) return false;
  const RsParams p = d_dgrad_params(c, l, 0, Bt, nullptr);
  return tc_rsgemm_supported(p) && tc_rsgemm2_supported(p) && (p.seg.nphase == 2 || p.merged_phases);
}
static int d_dgrad_layer(cg_ctx* c, int l, int b0, int nb, void* out, double* sumsq = nullptr, const void* ps_mask = nullptr,
                         int group_b = 0, const int32_t* sh = nullptr, int groups = 0) {
  RsParams p = d_dgrad_params(c, l, b0, nb, out);
  p.sumsq = sumsq;
  if (ps_mask) {
    p.epi = EPI_PS_MASK; p.mask = ps_mask;
    p.ps_w = c->dl[l - 1]; p.ps_group_b = group_b;
    for (int i = 0; i < 4; ++i) p.ps_shift[i] = i < groups ? sh[i * 4 + (l - 2)] : 0;
  }
  return launch_rsgemm(c, p);
}
static RsParams d_dgrad_params(cg_ctx* c, int l, int b0, int nb, void* out) {
  RsParams p;
  memset(&p, 0, sizeof(p));
  p.A = off(c, c->DA[l], (long long)b0 * c->dl[l] * c->dcp[l]);
  p.a_bs = (long long)c->dl[l] * c->dcp[l]; p.a_rs = c->dcp[l]; p.a_rows = c->dl[l];
  p.W = c->Wb_d[l]; p.w_ld = c->K * c->dcp[l];
  p.out = out; p.o_bs = (long long)c->dl[l - 1] * c->dcp[l - 1]; p.o_rs = 2 * c->dcp[l - 1]; p.o_phase_col = c->dcp[l - 1];
  p.B = nb; p.Q = c->dl[l]; p.N = c->dcp[l - 1]; p.n_real = c->dc[l - 1]; p.Kc = c->dcp[l]; p.k_real = c->dc[l]; p.epi = EPI_NONE;
  p.seg = seg_transposed(c->K, c->dcp[l]);
  if (c->Wb2_d[l] && !c->tc.force_v1) {
    // merged phases: one N = 128 GEMM, columns [phase 0 | phase 1] = the contiguous (B, L/2, 2*Cp) output row; input
    // window j (shift emin + j) carries phase 0's tap k = padL - 2e and phase 1's tap k = padL + 1 - 2e
    p.W = c->Wb2_d[l]; p.w_ld = c->mp_K2 * c->dcp[l];
    p.N = 2 * c->dcp[l - 1];
    p.merged_phases = 1;
    p.flop_scale = (float)c->K / (float)c->mp_K2;
    memset(&p.seg, 0, sizeof(p.seg));
    p.seg.nphase = 1;
    p.seg.nseg[0] = c->mp_K2;
    for (int j = 0; j < c->mp_K2; ++j) {
      p.seg.shift[0][j] = (short)(c->mp_emin + j);
      p.seg.acol[0][j] = 0;
      p.seg.wk[0][j] = j * c->dcp[l];
    }
  }
  return p;
}

// backward chain of sum_b coef[b]*D(x)_b down to DA[1] (and DX[0] for samples [dx0_b0, dx0_b0+dx0_nb))
static int d_backward(cg_ctx* c, int Bt, int B, int groups, const int32_t* sh, int dx0_b0, int dx0_nb,
                      double* sumsq = nullptr) {   // sumsq: per-sample squared norm of dX0, fused into the last GEMM
  CK(glue(c, 2.0 * Bt * c->dl[NL] * c->dcp[NL] * c->esz + 4.0 * c->dl[NL] * c->dc[NL]));
  DISPATCH_T(c, glue_launch(head_backward_kernel<T>, dim3(grid_for((long long)Bt * c->dl[NL] * c->dcp[NL] / (16 / c->esz))), dim3(256), 0, c->stream, 
                    (const T*)c->H[NL], dparam(c, 10), c->coef, (T*)c->DA[NL], Bt, c->dl[NL], c->dc[NL], c->dcp[NL]));
  CK(post_launch(c, "head_bwd"));
  for (int l = NL; l >= 2; --l) {
    if (d_dgrad_ps_fusable(c, l, Bt)) {   // data gradient + PhaseShuffle adjoint + LeakyReLU slope in one kernel
      CK(d_dgrad_layer(c, l, 0, Bt, c->DA[l - 1], nullptr, c->H[l - 1], B, sh, groups));
      continue;
    }
    CK(d_dgrad_layer(c, l, 0, Bt, c->DX[l - 1]));
    const long long tot = (long long)Bt * c->dl[l - 1] * c->dcp[l - 1] / (16 / c->esz);
    CK(glue(c, 3.0 * Bt * c->dl[l - 1] * c->dcp[l - 1] * c->esz));
    DISPATCH_T(c, glue_launch(ps_scatter_mask_kernel<T>, dim3(grid_for(tot)), dim3(256), 0, c->stream, 
                      (const T*)c->DX[l - 1], (const T*)c->H[l - 1], (T*)c->DA[l - 1], Bt, B, c->dl[l - 1],
                      c->dcp[l - 1], group_shifts(sh, groups, l - 1)));
    CK(post_launch(c, "ps_scatter_mask"));
  }
  if (dx0_nb > 0) CK(d_dgrad_layer(c, 1, dx0_b0, dx0_nb, c->DX[0], sumsq));
  return 0;
}

// dW[k][ci][co] = sum_{b,o} X[l-1][b, 2o + k - padL, ci] * DA[l][b, o, co]
static WgParams conv_wgrad_params(cg_ctx* c, int l, int Bt) {
  WgParams w;
  memset(&w, 0, sizeof(w));
  const SegTable st = seg_strided(c->K, c->dcp[l - 1]);
  w.S = c->X[l - 1]; w.s_bs = (long long)c->dl[l - 1] * c->dcp[l - 1]; w.s_rs = 2 * c->dcp[l - 1]; w.s_rows = c->dl[l - 1] / 2;
  w.P = c->DA[l]; w.p_bs = (long long)c->dl[l] * c->dcp[l]; w.p_rs = c->dcp[l];
  w.dW = dgrad(c, 2 * (l - 1)); w.m_real = c->dc[l - 1]; w.n_real = c->dc[l];
  w.B = Bt; w.Q = c->dl[l]; w.Mp = c->dcp[l - 1]; w.Np = c->dcp[l]; w.nseg = c->K;
  for (int k = 0; k < c->K; ++k) { w.shift[k] = st.shift[0][k]; w.scol[k] = st.acol[0][k]; }
  return w;
}

static bool side_ok(cg_ctx* c);
static int side_fork(cg_ctx* c);
static int side_join(cg_ctx* c);
// bias gradients (column sums of DA[1..5] over the first nb_bias samples) on the side stream, under the GEMMs that follow.
// Opt-in experiment (CG_SIDE_GLUE=1): measured same-box 12.05 / 12.12 ms per step with it against 11.80 / 11.93 without --
// the small CTAs do become resident beside the tensor-core CTAs, but they take issue slots and L2 bandwidth from kernels
// that are statically partitioned over the SMs, and the slowest SM sets each GEMM's time. The signal metrics (one launch
// per step, under the generator's backward pass) are the only by-product kept on the side stream.
static bool side_glue_ok(cg_ctx* c) {
  if (!side_ok(c) || !c->bf || !getenv("CG_SIDE_GLUE")) return false;
  for (int l = 1; l <= NL; ++l) if (c->dcp[l] / 8 > 128) return false;
  return true;
}
static int side_colsum(cg_ctx* c, int nb_bias) {
  ColsumOps ops;
  ops.n = NL;
  for (int l = 1; l <= NL; ++l) {
    ops.op[l - 1].X = c->DA[l]; ops.op[l - 1].out = dgrad(c, 2 * (l - 1) + 1);
    ops.op[l - 1].rows = (long long)nb_bias * c->dl[l]; ops.op[l - 1].Cp = c->dcp[l]; ops.op[l - 1].c_real = c->dc[l];
  }
  CK(side_fork(c));
  colsum_light_kernel<bf16><<<dim3(148, ops.n), 128, 0, c->side>>>(ops);
  return post_launch(c, "colsum_side");
}
static int side_head_wgrad(cg_ctx* c, int Bt, int nb_bias, int tail_from) {
  const int tot = c->dl[NL] * c->dcp[NL] / 8;
  dim3 hgrid(grid_for(tot, 128), Bt >= 64 ? 32 : 1);
  CK(side_fork(c));
  head_wgrad_kernel<bf16><<<hgrid, 128, 0, c->side>>>((const bf16*)c->X[NL], (const bf16*)c->V5, tail_from, c->coef, dgrad(c, 10),
                                                     dgrad(c, 11), Bt, nb_bias, c->dl[NL], c->dc[NL], c->dcp[NL]);
  return post_launch(c, "head_wgrad_side");
}

// all critic weight gradients from X[l-1] x DA[l] over Bt samples; biases from the first nb_bias samples.
// side_glue: the bias column sums and the head's weight gradient already run on the side stream (side_colsum /
// side_head_wgrad); they are joined after the first GEMM, before the first gradient bucket is declared complete
static int d_wgrad(cg_ctx* c, int Bt, int nb_bias, int tail_from, bool side_glue = false) {   // samples >= tail_from: head input = V5
  if (side_glue) {
    for (int l = NL; l >= 1; --l) {
      CK(launch_wgrad(c, conv_wgrad_params(c, l, Bt)));
      if (l == NL) { CK(side_join(c)); CU(cudaEventRecord(c->bucket_evt[CG_DISCRIMINATOR][0], c->stream)); }
      if (l == NL - 1) CU(cudaEventRecord(c->bucket_evt[CG_DISCRIMINATOR][1], c->stream));
    }
    CU(cudaEventRecord(c->bucket_evt[CG_DISCRIMINATOR][2], c->stream));
    return 0;
  }
  if (nb_bias > 0) {   // all five bias gradients in one launch
    ColsumOps ops;
    ops.n = NL;
    bool vec_ok = true;
    for (int l = 1; l <= NL; ++l) {
      ops.op[l - 1].X = c->DA[l]; ops.op[l - 1].out = dgrad(c, 2 * (l - 1) + 1);
      ops.op[l - 1].rows = (long long)nb_bias * c->dl[l]; ops.op[l - 1].Cp = c->dcp[l]; ops.op[l - 1].c_real = c->dc[l];
      if (c->dcp[l] / (16 / c->esz) > 256) vec_ok = false;
    }
    if (vec_ok) CK(launch_colsum_ops(c, ops));
    else
      for (int l = 1; l <= NL; ++l)
        CK(launch_colsum(c, ops.op[l - 1].X, ops.op[l - 1].out, ops.op[l - 1].rows, ops.op[l - 1].Cp, ops.op[l - 1].c_real));
  }
  const int tot = c->dl[NL] * c->dcp[NL] / (16 / c->esz);
  dim3 hgrid(grid_for(tot), Bt >= 64 ? 32 : 1);
  CK(glue(c, (double)Bt * c->dl[NL] * c->dcp[NL] * c->esz));
  DISPATCH_T(c, glue_launch(head_wgrad_kernel<T>, dim3(hgrid), dim3(256), 0, c->stream, 
                    (const T*)c->X[NL], (const T*)c->V5, tail_from, c->coef, dgrad(c, 10), dgrad(c, 11), Bt, nb_bias, c->dl[NL],
                    c->dc[NL], c->dcp[NL]));
  CK(post_launch(c, "head_wgrad"));
  // reverse layer order: the data-parallel host all-reduces bucket b as soon as its last writer has finished
  for (int l = NL; l >= 1; --l) {
    CK(launch_wgrad(c, conv_wgrad_params(c, l, Bt)));
    if (l == NL) CU(cudaEventRecord(c->bucket_evt[CG_DISCRIMINATOR][0], c->stream));
    if (l == NL - 1) CU(cudaEventRecord(c->bucket_evt[CG_DISCRIMINATOR][1], c->stream));
  }
  CU(cudaEventRecord(c->bucket_evt[CG_DISCRIMINATOR][2], c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------ rng
// `calls` consecutive draws of n floats each (out + call * n); draw number (first_call + call) of this rank's stream
static int draw(cg_ctx* c, float* out, long long n, int calls, uint64_t first_call, int mode) {
  const uint64_t stream0 = (first_call << 24) | ((uint64_t)(c->cfg.rank & 0xFFFFF) << 4) | (uint64_t)mode;
  dim3 grid(grid_for((n + 3) / 4), calls);
  rng_fill_kernel<<<grid, 256, 0, c->stream>>>(out, n, c->seed, stream0, mode);
  return post_launch(c, "rng_fill");
}
// shift number i of the run: identical on every rank (per-call scalars shared by the whole batch, calciumgan.py:121-124)
static int32_t shift_draw(const cg_ctx* c, uint64_t i) {
  const int m = c->cfg.phase_m;
  const uint64_t r = splitmix64(splitmix64(c->seed ^ 0x5048415345ull) + i);
  return (int32_t)(r % (uint64_t)(2 * m + 1)) - m;
}

static int check_batch(cg_ctx* c, int B) {
  if (B < 1 || B > c->Bmax) return set_err("batch %d outside [1, max_batch=%d]", B, c->Bmax);
  return 0;
}

// ------------------------------------------------------------------------------------------ Adam
static int apply_update_from(cg_ctx* c, int which, const float* grad);
extern "C" int cg_apply_update(cg_ctx* c, int which) { return apply_update_from(c, which, model_of(c, which)->g); }
extern "C" int cg_apply_update_reduced(cg_ctx* c, int which) {
  Model* m = model_of(c, which);
  if (!m->gr) return set_err("cg_apply_update_reduced: no reduced gradient (call cg_reduce_peer_grads first)");
  return apply_update_from(c, which, m->gr);
}
static int apply_update_from(cg_ctx* c, int which, const float* grad) {
  Model* m = model_of(c, which);
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-7f, gscale = 1.0f / (float)c->cfg.world_size;
  // pass 1 (reads the gradient once): non-finite check; the last block advances `iterations` and computes lr_t
  CK(glue(c, 4.0 * m->total));
  glue_launch(adam_prepare_kernel, dim3(grid_for(m->total / 4, 256, 148 * 4)), dim3(256), 0, c->stream, grad, m->total, m->opt,
                                                                                  c->cfg.learning_rate, b1, b2);
  CK(post_launch(c, "adam_prepare"));
  if (c->dbg_flags & CG_DEBUG_NO_ADAM_FUSE) {
    CK(glue(c, 28.0 * m->total));
    adam_kernel<<<grid_for(m->total), 256, 0, c->stream>>>(m->w, m->m, m->v, grad, m->total, m->opt, b1, b2, eps, gscale);
    CK(post_launch(c, "adam"));
    return repack(c, which);
  }
  const AdamPlan& pl = c->adam_plan[which];
  CK(glue(c, 28.0 * m->total + 2.0 * c->esz * m->total));
  DISPATCH_T(c, glue_launch(adam_pack_kernel<T>, dim3(grid_for(pl.items * 256, 256, 148 * 8)), dim3(256), 0, c->stream, 
                    m->w, m->m, m->v, grad, pl, m->opt, b1, b2, eps, gscale));
  return post_launch(c, "adam_pack");
}

static int fetch_scalars(cg_ctx* c, int slot, int flags, float* scalars_host) {
  if (flags & CG_FLAG_NO_SYNC) return 0;
  CU(cudaMemcpyAsync(c->h_scal, c->d_scal + (size_t)slot * CG_NUM_SCALARS, CG_NUM_SCALARS * 4, cudaMemcpyDeviceToHost,
                     c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (scalars_host) memcpy(scalars_host, c->h_scal, CG_NUM_SCALARS * 4);
  return 0;
}

// gan.py:32-41 / signals_metrics.py:9-28 on (real, FAKE32); acc[0..3] must be zeroed by the caller
static int launch_metrics(cg_ctx* c, const float* real, float* acc, long long rows, const float* fake = nullptr) {
  if (!fake) fake = c->FAKE32;
  CK(glue(c, 8.0 * rows * c->C));
  if (((reinterpret_cast<uintptr_t>(real) | reinterpret_cast<uintptr_t>(fake)) & 15) == 0 && c->C <= 256 && !getenv("CG_NO_METRICS_SMEM")) {
    const size_t smem = (size_t)2 * 32 * c->C * sizeof(float);
    if (smem > 48 * 1024)   // per device, so not cached in a process-wide static; once per step
      CU(cudaFuncSetAttribute(metrics_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const long long nchunks = (rows + 31) / 32;
    const int grid = (int)(nchunks < 148 * 8 ? nchunks : 148 * 8);
    metrics_smem_kernel<<<grid, 256, smem, c->stream>>>(real, fake, acc, rows, c->C, c->cfg.signals_min, c->cfg.signals_max,
                                                        c->cfg.normalize);
  } else if (c->C % 2 == 0 && c->C <= 128 && ((reinterpret_cast<uintptr_t>(real) | reinterpret_cast<uintptr_t>(fake)) & 7) == 0)
    metrics8_kernel<<<grid_for(rows * 8, 256, 148 * 8), 256, 0, c->stream>>>(real, fake, acc, rows, c->C, c->cfg.signals_min,
                                                                            c->cfg.signals_max, c->cfg.normalize);
  else
    metrics_kernel<<<grid_for(rows * 32, 256, 148 * 4), 256, 0, c->stream>>>(real, fake, acc, rows, c->C, c->cfg.signals_min,
                                                                            c->cfg.signals_max, c->cfg.normalize);
  return post_launch(c, "metrics");
}

// ------------------------------------------------------------------------------------------ critic step
// generator part of a critic sub-step (wgan_gp.py:65-66 + the interpolation of :38-41): reads no critic weight, so the
// data-parallel host may run it under the all-reduce of the previous critic update (cg_prefetch_generator)
static int critic_generator_part(cg_ctx* c, const float* real, int B, const float* noise, const float* alpha,
                                 bool want_fake32) {
  const long long per = (long long)B * c->L * c->dcp[0];
  // generator head writes fp32 FAKE32 and the compute-type copy straight into the critic's "fake" slot
  bool xhat_done = false;
  CK(g_forward(c, noise, B, off(c, c->X[0], per), false, off(c, c->X[0], 2 * per), real, alpha, want_fake32, &xhat_done));
  if (!xhat_done) {
    CK(glue(c, (double)B * c->L * (8.0 * c->C + (double)c->dcp[0] * c->esz)));
    DISPATCH_T(c, interp_kernel<T><<<grid_for(per / 4), 256, 0, c->stream>>>(real, c->FAKE32, alpha,
                                                                            (T*)off(c, c->X[0], 2 * per), B, c->L, c->C,
                                                                            c->dcp[0]));
    CK(post_launch(c, "interp"));
  }
  return 0;
}

// forward part shared by cg_critic_step and cg_validate: fake, D on [real; fake; xhat], dgrad chain, GP scalars
static int critic_forward_gp(cg_ctx* c, const float* real, int B, const float* noise, const float* alpha,
                             const int32_t* sh, int slot, bool real_ready = false, bool train = false,
                             bool want_fake32 = true, bool gen_done = false) {
  const long long per = (long long)B * c->L * c->dcp[0];
  if (train) CU(cudaMemsetAsync(c->dis.g, 0, c->dis.total * 4, c->stream));
  if (!gen_done) CK(critic_generator_part(c, real, B, noise, alpha, want_fake32));
  if (!real_ready) {   // the 5 critic sub-steps of one train step share the real batch (wgan_gp.py:85-86)
    CK(glue(c, (double)B * c->L * (4.0 * c->C + (double)c->dcp[0] * c->esz)));
    DISPATCH_T(c, glue_launch(assemble_x0_kernel<T>, dim3(grid_for(per / 4)), dim3(256), 0, c->stream, real, nullptr, nullptr, (T*)c->X[0], B,
                                                                                 c->L, c->C, c->dcp[0], 1));
    CK(post_launch(c, "real_to_x0"));
  }
  CK(d_forward(c, 3 * B, B, 3, sh));
  glue_launch(fill_coef_kernel, dim3((3 * B + 255) / 256), dim3(256), 0, c->stream, c->coef, B, 3, 0.f);
  CK(post_launch(c, "fill_coef"));
  CU(cudaMemsetAsync(c->sumsq, 0, (size_t)B * 8, c->stream));
  // tensor-core path: ||g_b||^2 accumulates in the epilogue of the last data-gradient GEMM (fp32 accumulators)
  const bool fuse_norm = c->use_tc && c->dl[1] >= 128 && !c->tc.force_v1;
  CK(d_backward(c, 3 * B, B, 3, sh, 2 * B, B, fuse_norm ? c->sumsq : nullptr));
  if (!fuse_norm) {
    const long long per_sample = (long long)c->L * c->dcp[0];
    const int chunks = 8;
    CK(glue(c, (double)B * per_sample * c->esz));
    DISPATCH_T(c, sumsq_kernel<T><<<B * chunks, 256, 0, c->stream>>>((const T*)c->DX[0], c->sumsq, per_sample, chunks));
    CK(post_launch(c, "sumsq"));
  }
  glue_launch(critic_scalars_kernel, dim3(1), dim3(256), 0, c->stream, c->scores, c->sumsq, c->ucoef, c->norms,
                                                  c->d_scal + (size_t)slot * CG_NUM_SCALARS, B, c->cfg.gp_lambda);
  return post_launch(c, "critic_scalars");
}

// GP second-order term without the second-order graph, passes 3 and 4 input: v0 = u = d(lambda*GP)/dg, then
// the linearised forward v_l = PS(M_l * conv(v_{l-1})) into the activation slots of sample group `xg` (the x_hat group:
// 2 inside a critic step, 0 in the stand-alone gradient-penalty entry); sh4 = that group's four shifts
static int gp_linearised_forward(cg_ctx* c, int B, int xg, const int32_t* sh4) {
  const long long per = (long long)c->L * c->dcp[0];
  CK(glue(c, 2.0 * B * per * c->esz));
  DISPATCH_T(c, glue_launch(scale_rows_kernel<T>, dim3(grid_for(per * B / (16 / c->esz))), dim3(256), 0, c->stream, 
                    (const T*)c->DX[0], c->ucoef, (T*)off(c, c->X[0], (long long)xg * B * per), per, per * B));
  CK(post_launch(c, "scale_rows"));
  for (int l = 1; l <= NL; ++l) {
    const long long gin = (long long)xg * B * c->dl[l - 1] * c->dcp[l - 1], gout = (long long)xg * B * c->dl[l] * c->dcp[l];
    void* dst = l < NL ? off(c, c->DX[l], gout) : c->V5;   // H[l] of the x_hat group stays intact (slope masks, debug taps)
    RsParams p = conv_fwd_params(c, l, off(c, c->X[l - 1], gin), dst, B, EPI_MASK, off(c, c->H[l], gout));
    if (l < NL && ps_fusable(c, p)) {   // v_l = PS(M_l * conv(v_{l-1})) written straight into the xhat group's X_l slot
      p.out = nullptr;
      set_ps(p, off(c, c->X[l], gout), c->dl[l], B, sh4, 1, l);
      CK(launch_rsgemm(c, p));
      continue;
    }
    CK(launch_rsgemm(c, p));
    if (l < NL) {
      const long long tot = (long long)B * c->dl[l] * c->dcp[l] / (16 / c->esz);
      GroupShifts g; g.s[0] = sh4[l - 1]; g.s[1] = g.s[2] = g.s[3] = 0;
      DISPATCH_T(c, ps_gather_kernel<T><<<grid_for(tot), 256, 0, c->stream>>>(
                        (const T*)dst, (T*)off(c, c->X[l], gout), B, B, c->dl[l], c->dcp[l], g));
      CK(post_launch(c, "ps_gather_lin"));
    }
  }
  return 0;
}

static int critic_step_impl(cg_ctx* c, const float* real, int B, const float* noise, const float* alpha,
                            const int32_t* sh, int flags, int slot, bool real_ready = false, bool want_fake32 = true) {
  CK(critic_forward_gp(c, real, B, noise, alpha, sh, slot, real_ready, true, want_fake32,
                       (flags & CG_FLAG_GEN_PREFETCHED) != 0));
  // memory-bound by-products with no consumer before Adam go to the side stream, where their small CTAs are resident
  // beside the tensor-core CTAs of the gradient-penalty passes: bias gradients now, the head's weight gradient once v_5 exists
  const bool side = side_glue_ok(c);
  if (side) CK(side_colsum(c, 2 * B));
  CK(gp_linearised_forward(c, B, 2, sh + 8));
  if (side) CK(side_head_wgrad(c, 3 * B, 2 * B, 2 * B));
  CK(d_wgrad(c, 3 * B, 2 * B, 2 * B, side));
  if (!(flags & CG_FLAG_NO_UPDATE)) CK(cg_apply_update(c, CG_DISCRIMINATOR));
  return 0;
}

// (prep_random) Random inputs of one step function: `noise_calls` draws of (B, nd) noise, `alpha_calls` draws of (B) alpha and nsh
// PhaseShuffle shifts. Injected values are used as given; the stream positions advance by the same amount either way,
// so a rank that injects and a rank that does not stay in step, and the shared shift stream never depends on what a
// caller injected (two ranks making the same sequence of calls always see the same shifts).
static int prep_noise_alpha(cg_ctx* c, int B, const float*& noise, int noise_calls, const float** alpha, int alpha_calls) {
  if (!noise && noise_calls > 0) {
    CK(draw(c, c->noise_buf, (long long)B * c->nd, noise_calls, c->noise_calls, 0));
    noise = c->noise_buf;
  }
  c->noise_calls += (uint64_t)noise_calls;
  if (alpha && !*alpha && alpha_calls > 0) {
    CK(draw(c, c->alpha_buf, B, alpha_calls, c->alpha_calls, 1));
    *alpha = c->alpha_buf;
  }
  c->alpha_calls += (uint64_t)alpha_calls;
  c->last_noise = noise; c->last_n_noise = (int64_t)noise_calls * B * c->nd;
  c->last_alpha = alpha ? *alpha : nullptr; c->last_n_alpha = alpha ? (int64_t)alpha_calls * B : 0;
  return 0;
}
static int prep_shifts(cg_ctx* c, const int32_t*& sh, int32_t* shbuf, int nsh) {
  if (!sh) {
    for (int i = 0; i < nsh; ++i) shbuf[i] = shift_draw(c, c->shift_draws + (uint64_t)i);
    sh = shbuf;
  }
  c->shift_draws += (uint64_t)nsh;
  for (int i = 0; i < nsh; ++i)
    if (sh[i] < -c->cfg.phase_m || sh[i] > c->cfg.phase_m)
      return set_err("phase-shuffle shift %d outside [-m, m] (m=%d)", sh[i], c->cfg.phase_m);
  c->last_shifts.assign(sh, sh + nsh);
  return 0;
}
static int prep_random(cg_ctx* c, int B, const float*& noise, int noise_calls, const float** alpha, int alpha_calls,
                       const int32_t*& sh, int32_t* shbuf, int nsh) {
  CK(prep_noise_alpha(c, B, noise, noise_calls, alpha, alpha_calls));
  return prep_shifts(c, sh, shbuf, nsh);
}
// a step that follows cg_prefetch_generator: noise / alpha were drawn (and the streams advanced) there
static int take_prefetched(cg_ctx* c, int B, int for_gen_step, const float*& noise, const float** alpha) {
  if (!c->pref.valid || c->pref.for_gen_step != for_gen_step || c->pref.B != B)
    return set_err("CG_FLAG_GEN_PREFETCHED without a matching cg_prefetch_generator call");
  noise = c->pref.noise;
  if (alpha) *alpha = c->pref.alpha;
  c->pref.valid = false;
  c->last_noise = noise; c->last_n_noise = (int64_t)B * c->nd;
  c->last_alpha = alpha ? *alpha : nullptr; c->last_n_alpha = alpha ? B : 0;
  return 0;
}

extern "C" int cg_critic_step(cg_ctx* c, const float* real, int B, const float* noise, const float* alpha,
                              const int32_t* sh, int flags, float* scalars_host) {
  CK(check_batch(c, B));
  int32_t shbuf[12];
  if (flags & CG_FLAG_GEN_PREFETCHED) {
    CK(take_prefetched(c, B, 0, noise, &alpha));
    CK(prep_shifts(c, sh, shbuf, 12));
  } else {
    CK(prep_random(c, B, noise, 1, &alpha, 1, sh, shbuf, 12));
  }
  CK(critic_step_impl(c, real, B, noise, alpha, sh, flags, 0, (flags & CG_FLAG_SAME_REAL) != 0, !(flags & CG_FLAG_NO_FAKE32)));
  return fetch_scalars(c, 0, flags, scalars_host);
}

// gan.py:32-41 on a side stream, forked from the main stream now (FAKE32 is final) and joined by side_metrics_join: the
// metrics have no consumer before the scalars are read, and the kernel (no shared memory, 128-thread blocks) is small
// enough to be resident beside the tensor-core CTAs of the backward pass instead of taking its own 70 us of the step
static bool side_ok(cg_ctx* c) { return c->use_tc && !c->profiling && !getenv("CG_NO_SIDE_STREAM"); }
static bool side_metrics_ok(cg_ctx* c, const float* real) {
  return side_ok(c) && !getenv("CG_NO_SIDE_METRICS") && c->C % 2 == 0 && c->C <= 128 &&
         ((reinterpret_cast<uintptr_t>(real) | reinterpret_cast<uintptr_t>(c->FAKE32)) & 7) == 0;
}
// side stream starts after everything enqueued on the main stream so far / main stream waits for the side stream
static int side_fork(cg_ctx* c) {
  if (!c->side) {
    CU(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  }
  CU(cudaEventRecord(c->ev_fork, c->stream));
  CU(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
  return 0;
}
static int side_join(cg_ctx* c) {
  CU(cudaEventRecord(c->ev_join, c->side));
  CU(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  return 0;
}
static int side_metrics_fork(cg_ctx* c, const float* real, float* acc, long long rows) {
  CU(cudaMemsetAsync(acc, 0, 4 * 4, c->stream));
  CK(side_fork(c));
  metrics8_kernel<<<grid_for(rows * 8, 128, 148 * 4), 128, 0, c->side>>>(real, c->FAKE32, acc, rows, c->C, c->cfg.signals_min,
                                                                       c->cfg.signals_max, c->cfg.normalize);
  return post_launch(c, "metrics_side");
}
static int side_metrics_join(cg_ctx* c) { return side_join(c); }

// ------------------------------------------------------------------------------------------ generator step
// generator forward of the generator step (wgan_gp.py:23-26): keeps everything its backward pass needs
static int generator_step_generator_part(cg_ctx* c, int B, const float* noise) {
  CU(cudaMemcpyAsync(c->Z, noise, (size_t)B * c->nd * 4, cudaMemcpyDeviceToDevice, c->stream));
  return g_forward(c, c->Z, B, c->X[0]);
}
static int generator_step_impl(cg_ctx* c, const float* real, int B, const float* noise, const int32_t* sh, int flags,
                               int slot) {
  if (!(flags & CG_FLAG_GEN_PREFETCHED)) CK(generator_step_generator_part(c, B, noise));
  float* scal = c->d_scal + (size_t)slot * CG_NUM_SCALARS;
  const bool side = real && side_metrics_ok(c, real);
  if (side) CK(side_metrics_fork(c, real, scal + CG_S_MET_MIN, (long long)B * c->L));
  CK(d_forward(c, B, B, 1, sh));
  glue_launch(gen_loss_kernel, dim3(1), dim3(256), 0, c->stream, c->scores, scal, B);
  CK(post_launch(c, "gen_loss"));
  glue_launch(fill_coef_kernel, dim3((B + 255) / 256), dim3(256), 0, c->stream, c->coef, B, 1, -1.f / B);
  CK(post_launch(c, "fill_coef"));
  CK(d_backward(c, B, B, 1, sh, 0, B));
  CU(cudaMemsetAsync(c->gen.g, 0, c->gen.total * 4, c->stream));
  CK(g_backward(c, B));
  if (side) CK(side_metrics_join(c));
  else if (real) {
    CU(cudaMemsetAsync(scal + CG_S_MET_MIN, 0, 4 * 4, c->stream));
    const long long rows = (long long)B * c->L;
    CK(launch_metrics(c, real, scal + CG_S_MET_MIN, rows));
  }
  if (!(flags & CG_FLAG_NO_UPDATE)) CK(cg_apply_update(c, CG_GENERATOR));
  return 0;
}

extern "C" int cg_generator_step(cg_ctx* c, const float* real, int B, const float* noise, const int32_t* sh, int flags,
                                 float* scalars_host) {
  CK(check_batch(c, B));
  int32_t shbuf[4];
  if (flags & CG_FLAG_GEN_PREFETCHED) {
    CK(take_prefetched(c, B, 1, noise, nullptr));
    CK(prep_shifts(c, sh, shbuf, 4));
  } else {
    CK(prep_random(c, B, noise, 1, nullptr, 0, sh, shbuf, 4));
  }
  CK(generator_step_impl(c, real, B, noise, sh, flags, 0));
  return fetch_scalars(c, 0, flags, scalars_host);
}

// Data-parallel overlap (no reference counterpart): the generator forward of the NEXT sub-step reads no critic
// weight, so the host enqueues it before it waits for the all-reduce of the critic gradients of the current one.
extern "C" int cg_prefetch_generator(cg_ctx* c, const float* real, int B, const float* noise, const float* alpha,
                                     int for_generator_step, int flags) {
  CK(check_batch(c, B));
  if (c->pref.valid) return set_err("cg_prefetch_generator: the previous prefetch has not been consumed");
  if (for_generator_step) {
    CK(prep_noise_alpha(c, B, noise, 1, nullptr, 0));
    CK(generator_step_generator_part(c, B, noise));
  } else {
    if (!real) return set_err("cg_prefetch_generator: the critic sub-step needs the real batch (interpolation)");
    CK(prep_noise_alpha(c, B, noise, 1, &alpha, 1));
    CK(critic_generator_part(c, real, B, noise, alpha, !(flags & CG_FLAG_NO_FAKE32)));
  }
  c->pref.valid = true; c->pref.for_gen_step = for_generator_step ? 1 : 0; c->pref.B = B;
  c->pref.noise = noise; c->pref.alpha = alpha;
  return 0;
}

// wgan_gp.py:82-95
extern "C" int cg_train_step(cg_ctx* c, const float* real, int B, const float* noise, const float* alpha,
                             const int32_t* sh, float* scalars_host) {
  CK(check_batch(c, B));
  const int nc = c->cfg.n_critic;
  std::vector<int32_t> shbuf(12 * nc + 4);
  // same stream positions as nc cg_critic_step calls followed by one cg_generator_step (the data-parallel host path)
  CK(prep_random(c, B, noise, nc + 1, &alpha, nc, sh, shbuf.data(), 12 * nc + 4));
  for (int i = 0; i < nc; ++i)
    CK(critic_step_impl(c, real, B, noise + (size_t)i * B * c->nd, alpha + (size_t)i * B, sh + 12 * i, 0, i, i > 0,
                        /*want_fake32=*/false));   // the generator step rewrites FAKE32 before anything reads it
  CK(generator_step_impl(c, real, B, noise + (size_t)nc * B * c->nd, sh + 12 * nc, 0, nc));
  CU(cudaMemcpyAsync(c->h_scal, c->d_scal, (size_t)(nc + 1) * CG_NUM_SCALARS * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (scalars_host) {
    memset(scalars_host, 0, CG_NUM_SCALARS * 4);
    double dl = 0, gp = 0;
    for (int i = 0; i < nc; ++i) { dl += c->h_scal[i * CG_NUM_SCALARS + CG_S_DIS_LOSS]; gp += c->h_scal[i * CG_NUM_SCALARS + CG_S_GP]; }
    scalars_host[CG_S_DIS_LOSS] = (float)(dl / nc);
    scalars_host[CG_S_GP] = (float)(gp / nc);
    const float* g = c->h_scal + (size_t)nc * CG_NUM_SCALARS;
    for (int i = CG_S_GEN_LOSS; i <= CG_S_MET_STD; ++i) scalars_host[i] = g[i];
  }
  return 0;
}

// gan.py:58-70,87-90 with the WGAN-GP losses; no parameter update
extern "C" int cg_validate(cg_ctx* c, const float* real, int B, const float* noise, const float* alpha,
                           const int32_t* sh, float* fake_out, float* scalars_host) {
  CK(check_batch(c, B));
  int32_t shbuf[12];
  CK(prep_random(c, B, noise, 1, &alpha, 1, sh, shbuf, 12));
  CK(critic_forward_gp(c, real, B, noise, alpha, sh, 0));
  float* scal = c->d_scal;
  glue_launch(gen_loss_kernel, dim3(1), dim3(256), 0, c->stream, c->scores + B, scal, B);   // -mean D(fake)
  CK(post_launch(c, "gen_loss"));
  CU(cudaMemsetAsync(scal + CG_S_MET_MIN, 0, 4 * 4, c->stream));
  const long long rows = (long long)B * c->L;
  CK(launch_metrics(c, real, scal + CG_S_MET_MIN, rows));
  if (fake_out) CU(cudaMemcpyAsync(fake_out, c->FAKE32, (size_t)rows * c->C * 4, cudaMemcpyDeviceToDevice, c->stream));
  return fetch_scalars(c, 0, 0, scalars_host);
}

// batch assembly from a device-resident dataset cache: dst[i] = src[idx[i]] (rows of row_elems floats)
extern "C" int cg_gather_rows(cg_ctx* c, const float* src, int64_t n_src, const int64_t* idx_dev, int n, int64_t row_elems,
                              float* dst) {
  if (!src || !idx_dev || !dst || n < 1 || row_elems < 1 || n_src < 1) return set_err("cg_gather_rows: bad arguments");
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) || (row_elems & 3))
    return set_err("cg_gather_rows: rows must be 16-byte aligned (row_elems %% 4 == 0)");
  int gx = (int)((row_elems / 4 + 255) / 256);
  const int cap = (148 * 16 + n - 1) / n;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid(gx, n);
  CK(glue(c, 8.0 * n * row_elems));
  gather_rows_kernel<<<grid, 256, 0, c->stream>>>(src, (const long long*)idx_dev, dst, row_elems, n_src);
  return post_launch(c, "gather_rows");
}

// gan.py:32-41 on caller-provided tensors: out_host[4] = min, max, mean, std errors (signals_metrics.py:9-28)
extern "C" int cg_metrics(cg_ctx* c, const float* real, const float* fake, int B, float* out_host) {
  if (!real || !fake || !out_host || B < 1) return set_err("cg_metrics: bad arguments");
  float* scal = c->d_scal;
  CU(cudaMemsetAsync(scal + CG_S_MET_MIN, 0, 4 * 4, c->stream));
  CK(launch_metrics(c, real, scal + CG_S_MET_MIN, (long long)B * c->L, fake));
  CU(cudaMemcpyAsync(c->h_scal, scal, CG_NUM_SCALARS * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 4; ++i) out_host[i] = c->h_scal[CG_S_MET_MIN + i];
  return 0;
}

extern "C" int cg_generate(cg_ctx* c, const float* noise, int B, int denorm, float* out) {
  CK(check_batch(c, B));
  if (!noise || !out) return set_err("cg_generate: null pointer");
  CK(g_forward(c, noise, B));
  const long long tot = (long long)B * c->L * c->C;
  if (denorm) {
    denorm_kernel<<<grid_for(tot), 256, 0, c->stream>>>(c->FAKE32, out, tot, c->cfg.signals_min, c->cfg.signals_max);
    CK(post_launch(c, "denorm"));
  } else {
    CU(cudaMemcpyAsync(out, c->FAKE32, (size_t)tot * 4, cudaMemcpyDeviceToDevice, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------ debug taps
extern "C" int cg_debug_critic_forward(cg_ctx* c, const float* x, int B, const int32_t* sh, float* scores_dev) {
  CK(check_batch(c, B));
  if (!x || !sh || !scores_dev) return set_err("cg_debug_critic_forward: null pointer");
  const long long tot = (long long)B * c->L * c->dcp[0] / 4;
  DISPATCH_T(c, glue_launch(assemble_x0_kernel<T>, dim3(grid_for(tot)), dim3(256), 0, c->stream, x, nullptr, nullptr, (T*)c->X[0], B, c->L,
                                                                           c->C, c->dcp[0], 1));
  CK(post_launch(c, "assemble_x0"));
  CK(d_forward(c, B, B, 1, sh));
  CU(cudaMemcpyAsync(scores_dev, c->scores, (size_t)B * 4, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int cg_debug_gp(cg_ctx* c, const float* xhat, int B, const int32_t* sh, float* grad_dev, float* norms_dev) {
  CK(check_batch(c, B));
  if (!xhat || !sh) return set_err("cg_debug_gp: null pointer");
  const long long tot = (long long)B * c->L * c->dcp[0] / 4;
  DISPATCH_T(c, glue_launch(assemble_x0_kernel<T>, dim3(grid_for(tot)), dim3(256), 0, c->stream, xhat, nullptr, nullptr, (T*)c->X[0], B,
                                                                           c->L, c->C, c->dcp[0], 1));
  CK(post_launch(c, "assemble_x0"));
  CK(d_forward(c, B, B, 1, sh));
  glue_launch(fill_coef_kernel, dim3((B + 255) / 256), dim3(256), 0, c->stream, c->coef, B, 1, 1.f);
  CK(post_launch(c, "fill_coef"));
  CK(d_backward(c, B, B, 1, sh, 0, B));
  if (grad_dev) {
    DISPATCH_T(c, unpad_kernel<T><<<grid_for((long long)B * c->L * c->C), 256, 0, c->stream>>>(
                      (const T*)c->DX[0], grad_dev, (long long)B * c->L, c->C, c->dcp[0]));
    CK(post_launch(c, "unpad"));
  }
  if (norms_dev) {
    CU(cudaMemsetAsync(c->sumsq, 0, (size_t)B * 8, c->stream));
    DISPATCH_T(c, sumsq_kernel<T><<<B * 8, 256, 0, c->stream>>>((const T*)c->DX[0], c->sumsq, (long long)c->L * c->dcp[0], 8));
    CK(post_launch(c, "sumsq"));
    sumsq_to_float_kernel<<<(B + 255) / 256, 256, 0, c->stream>>>(c->sumsq, norms_dev, B);   // squared norms
    CK(post_launch(c, "sumsq_to_float"));
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// BASELINE.json configs[4]: the gradient penalty alone -- critic forward at x_hat, data-gradient chain to g = dD/dx_hat,
// GP = mean (||g|| - 1)^2, and its gradient w.r.t. every critic weight through the double backward (linearised forward
// + weight gradients), i.e. all four passes of wgan_gp.py:43-50 under the outer tape of optimizer.py:32. Leaves
// lambda * dGP/dW in the critic's gradient buffer (cg_get_grads) and the GP value in scalars_host[CG_S_GP].
extern "C" int cg_gp_gradient(cg_ctx* c, const float* xhat, int B, const int32_t* sh, int flags, float* scalars_host) {
  CK(check_batch(c, B));
  if (!xhat || !sh) return set_err("cg_gp_gradient: null pointer");
  for (int i = 0; i < 4; ++i)
    if (sh[i] < -c->cfg.phase_m || sh[i] > c->cfg.phase_m) return set_err("cg_gp_gradient: shift outside [-m, m]");
  CU(cudaMemsetAsync(c->dis.g, 0, c->dis.total * 4, c->stream));
  const long long tot = (long long)B * c->L * c->dcp[0] / 4;
  DISPATCH_T(c, glue_launch(assemble_x0_kernel<T>, dim3(grid_for(tot)), dim3(256), 0, c->stream, xhat, nullptr, nullptr, (T*)c->X[0], B,
                                                                           c->L, c->C, c->dcp[0], 1));
  CK(post_launch(c, "assemble_x0"));
  CK(d_forward(c, B, B, 1, sh));                                                    // pass 1: forward, slope masks
  glue_launch(fill_coef_kernel, dim3((B + 255) / 256), dim3(256), 0, c->stream, c->coef, B, 1, 1.f);
  CK(post_launch(c, "fill_coef"));
  CU(cudaMemsetAsync(c->sumsq, 0, (size_t)B * 8, c->stream));
  const bool fuse_norm = c->use_tc && c->dl[1] >= 128 && !c->tc.force_v1;
  CK(d_backward(c, B, B, 1, sh, 0, B, fuse_norm ? c->sumsq : nullptr));             // pass 2: g and ||g||
  if (!fuse_norm) {
    DISPATCH_T(c, sumsq_kernel<T><<<B * 8, 256, 0, c->stream>>>((const T*)c->DX[0], c->sumsq, (long long)c->L * c->dcp[0], 8));
    CK(post_launch(c, "sumsq"));
  }
  glue_launch(critic_scalars_kernel, dim3(1), dim3(256), 0, c->stream, c->scores, c->sumsq, c->ucoef, c->norms, c->d_scal, B, c->cfg.gp_lambda);
  CK(post_launch(c, "critic_scalars"));
  CK(gp_linearised_forward(c, B, 0, sh));                                           // pass 3
  CK(d_wgrad(c, B, 0, 0));                                                          // pass 4 (dGP/db = 0)
  return fetch_scalars(c, 0, flags, scalars_host);
}

static int pad_in(cg_ctx* c, const float* src, void* dst, int B, int rows, int C, int Cp) {
  DISPATCH_T(c, glue_launch(assemble_x0_kernel<T>, dim3(grid_for((long long)B * rows * Cp / 4)), dim3(256), 0, c->stream, 
                    src, nullptr, nullptr, (T*)dst, B, rows, C, Cp, 1));
  return post_launch(c, "pad_in");
}
static int pad_out(cg_ctx* c, const void* src, float* dst, long long rows, int C, int Cp) {
  DISPATCH_T(c, unpad_kernel<T><<<grid_for(rows * C), 256, 0, c->stream>>>((const T*)src, dst, rows, C, Cp));
  return post_launch(c, "pad_out");
}

extern "C" int cg_debug_layer(cg_ctx* c, int which, int layer, int pass, const float* x, const float* dy, int B,
                              float* out) {
  if (layer < 1 || layer > NL || pass < 0 || pass > 2 || !out) return set_err("cg_debug_layer: bad arguments");
  if ((pass != 1 && !x) || (pass != 0 && !dy)) return set_err("cg_debug_layer: missing input");
  if (which == CG_DISCRIMINATOR) {
    const int l = layer;
    if (B < 1 || B > 3 * c->Bmax || (pass == 1 && l == 1 && B > c->Bmax)) return set_err("cg_debug_layer: batch too large");
    if (x) CK(pad_in(c, x, c->X[l - 1], B, c->dl[l - 1], c->dc[l - 1], c->dcp[l - 1]));
    if (dy) CK(pad_in(c, dy, c->DA[l], B, c->dl[l], c->dc[l], c->dcp[l]));
    if (pass == 0) {
      CK(launch_rsgemm(c, conv_fwd_params(c, l, c->X[l - 1], c->H[l], B, EPI_BIAS_LRELU, nullptr)));
      CK(pad_out(c, c->H[l], out, (long long)B * c->dl[l], c->dc[l], c->dcp[l]));
    } else if (pass == 1) {
      void* dst = l == 1 ? c->DX[0] : c->DX[l - 1];
      CK(d_dgrad_layer(c, l, 0, B, dst));
      CK(pad_out(c, dst, out, (long long)B * c->dl[l - 1], c->dc[l - 1], c->dcp[l - 1]));
    } else {
      const ParamInfo& pi = c->dis.params[2 * (l - 1)];
      CU(cudaMemsetAsync(c->dis.g + pi.offset, 0, pi.size * 4, c->stream));
      CK(launch_wgrad(c, conv_wgrad_params(c, l, B)));
      CU(cudaMemcpyAsync(out, c->dis.g + pi.offset, pi.size * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
  } else {
    const int i = layer;
    CK(check_batch(c, B));
    if (x) CK(pad_in(c, x, c->HG[i - 1], B, c->gl[i - 1], c->gc[i - 1], c->gcp[i - 1]));
    if (dy) CK(pad_in(c, dy, c->DAG[i], B, c->gl[i], c->gc[i], c->gcp[i]));
    if (pass == 0) {
      CK(launch_rsgemm(c, convT_fwd_params(c, i, B)));
      CK(pad_out(c, c->AG[i], out, (long long)B * c->gl[i], c->gc[i], c->gcp[i]));
    } else if (pass == 1) {
      CK(launch_rsgemm(c, convT_bwd_params(c, i, B)));
      CK(pad_out(c, c->DHG[i - 1], out, (long long)B * c->gl[i - 1], c->gc[i - 1], c->gcp[i - 1]));
    } else {
      const ParamInfo& pi = c->gen.params[c->g_k[i]];
      CU(cudaMemsetAsync(c->gen.g + pi.offset, 0, pi.size * 4, c->stream));
      CK(launch_wgrad(c, convT_wgrad_params(c, i, B)));
      CU(cudaMemcpyAsync(out, c->gen.g + pi.offset, pi.size * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int cg_phase_shuffle_index(int w, int shift, int32_t* idx) {
  if (w < 1 || !idx) return set_err("cg_phase_shuffle_index: bad arguments");
  if (shift > w - 1 || shift < -(w - 1)) return set_err("cg_phase_shuffle_index: |shift| must be < w");
  for (int t = 0; t < w; ++t) idx[t] = ps_index(t, shift, w);
  return 0;
}

extern "C" int cg_phase_shuffle_scatter_index(int w, int shift, int32_t* t1, int32_t* t2) {
  if (w < 1 || !t1 || !t2) return set_err("cg_phase_shuffle_scatter_index: bad arguments");
  if (shift > w - 1 || shift < -(w - 1)) return set_err("cg_phase_shuffle_scatter_index: |shift| must be < w");
  for (int q = 0; q < w; ++q) { int a, b; ps_scatter_targets(q, shift, w, a, b); t1[q] = a; t2[q] = b; }
  return 0;
}
extern "C" int cg_phase_shuffle_adjoint_plan(int w, int shift, int32_t* dest, int32_t* src_slot, int32_t* par_slot,
                                             int32_t* zero) {
  if (w < 2 || (w & 1) || !dest || !src_slot || !par_slot || !zero) return set_err("cg_phase_shuffle_adjoint_plan: bad arguments");
  if (shift > w - 1 || shift < -(w - 1)) return set_err("cg_phase_shuffle_adjoint_plan: |shift| must be < w");
  for (int t = 0; t < w; ++t) {
    int d, xs, xp; bool xz;
    ps_adjoint_row(t, shift, w, d, xs, xp, xz);
    dest[t] = d; src_slot[t] = xs; par_slot[t] = xp; zero[t] = xz ? 1 : 0;
  }
  return 0;
}

extern "C" int cg_debug_phase_shuffle(cg_ctx* c, const float* x, int B, int w, int ch, int shift, float* out) {
  if (!x || !out || B < 1 || w < 1 || ch < 4 || ch % 4) return set_err("cg_debug_phase_shuffle: bad arguments");
  if (shift > w - 1 || shift < -(w - 1)) return set_err("cg_debug_phase_shuffle: |shift| must be < w");
  GroupShifts g; g.s[0] = shift; g.s[1] = g.s[2] = g.s[3] = 0;
  ps_gather_kernel<float><<<grid_for((long long)B * w * ch / 4), 256, 0, c->stream>>>(x, out, B, B, w, ch, g);
  CK(post_launch(c, "ps_gather_debug"));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

static int debug_buffer(cg_ctx* c, int buffer, int layer, const void** ptr, int64_t* rows, int* ch, int* chp, int* cap) {
  const bool critic = buffer == CG_BUF_X || buffer == CG_BUF_H || buffer == CG_BUF_DA;
  const int lo = (buffer == CG_BUF_X || buffer == CG_BUF_HG) ? 0 : 1;
  if (layer < lo || layer > NL) return set_err("cg_debug_read: layer %d out of range for buffer %d", layer, buffer);
  switch (buffer) {
    case CG_BUF_X: *ptr = c->X[layer]; break;
    case CG_BUF_H: *ptr = c->H[layer]; break;
    case CG_BUF_DA: *ptr = c->DA[layer]; break;
    case CG_BUF_HG: *ptr = c->HG[layer]; break;
    case CG_BUF_AG: *ptr = c->AG[layer]; break;
    case CG_BUF_DAG: *ptr = c->DAG[layer]; break;
    default: return set_err("cg_debug_read: unknown buffer %d", buffer);
  }
  if (critic) { *rows = c->dl[layer]; *ch = c->dc[layer]; *chp = c->dcp[layer]; *cap = 3 * c->Bmax; }
  else { *rows = c->gl[layer]; *ch = c->gc[layer]; *chp = c->gcp[layer]; *cap = c->Bmax; }
  return 0;
}
extern "C" int cg_debug_buffer_shape(cg_ctx* c, int buffer, int layer, int64_t* rows, int64_t* channels) {
  const void* ptr = nullptr; int64_t r = 0; int ch = 0, chp = 0, cap = 0;
  CK(debug_buffer(c, buffer, layer, &ptr, &r, &ch, &chp, &cap));
  if (rows) *rows = r;
  if (channels) *channels = ch;
  return 0;
}
extern "C" int cg_debug_read(cg_ctx* c, int buffer, int layer, int batch, float* out) {
  const void* ptr = nullptr; int64_t rows = 0; int ch = 0, chp = 0, cap = 0;
  CK(debug_buffer(c, buffer, layer, &ptr, &rows, &ch, &chp, &cap));
  if (!out || batch < 1 || batch > cap) return set_err("cg_debug_read: bad batch %d (capacity %d)", batch, cap);
  CK(pad_out(c, ptr, out, (long long)batch * rows, ch, chp));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int cg_debug_last_draws(cg_ctx* c, float* noise_out, int64_t n_noise, float* alpha_out, int64_t n_alpha,
                                   int32_t* shifts_out, int n_shifts) {
  if (noise_out) {
    if (!c->last_noise || n_noise > c->last_n_noise) return set_err("cg_debug_last_draws: the last step used %lld noise values", (long long)c->last_n_noise);
    CU(cudaMemcpyAsync(noise_out, c->last_noise, (size_t)n_noise * 4, cudaMemcpyDeviceToDevice, c->stream));
  }
  if (alpha_out) {
    if (!c->last_alpha || n_alpha > c->last_n_alpha) return set_err("cg_debug_last_draws: the last step used %lld alpha values", (long long)c->last_n_alpha);
    CU(cudaMemcpyAsync(alpha_out, c->last_alpha, (size_t)n_alpha * 4, cudaMemcpyDeviceToDevice, c->stream));
  }
  if (shifts_out) {
    if (n_shifts > (int)c->last_shifts.size()) return set_err("cg_debug_last_draws: the last step used %d shifts", (int)c->last_shifts.size());
    memcpy(shifts_out, c->last_shifts.data(), (size_t)n_shifts * 4);
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// DA[l] -> DA[l-1]: data gradient, PhaseShuffle adjoint, LeakyReLU slope (one link of d_backward, in isolation)
extern "C" int cg_debug_dgrad_ps(cg_ctx* c, int layer, const float* dy, const float* h, int B, int group_b,
                                 const int32_t* shifts, float* out) {
  const int l = layer;
  if (l < 2 || l > NL || !dy || !h || !shifts || !out) return set_err("cg_debug_dgrad_ps: bad arguments");
  if (B < 1 || B > 3 * c->Bmax || group_b < 1) return set_err("cg_debug_dgrad_ps: batch out of range");
  const int groups = (B + group_b - 1) / group_b;
  if (groups > 3) return set_err("cg_debug_dgrad_ps: at most 3 shift groups");
  int32_t sh[12] = {0};
  for (int g = 0; g < groups; ++g) {
    if (shifts[g] < -c->cfg.phase_m || shifts[g] > c->cfg.phase_m) return set_err("cg_debug_dgrad_ps: shift outside [-m, m]");
    sh[g * 4 + (l - 2)] = shifts[g];
  }
  CK(pad_in(c, dy, c->DA[l], B, c->dl[l], c->dc[l], c->dcp[l]));
  CK(pad_in(c, h, c->H[l - 1], B, c->dl[l - 1], c->dc[l - 1], c->dcp[l - 1]));
  if (d_dgrad_ps_fusable(c, l, B)) {
    CK(d_dgrad_layer(c, l, 0, B, c->DA[l - 1], nullptr, c->H[l - 1], group_b, sh, groups));
  } else {
    CK(d_dgrad_layer(c, l, 0, B, c->DX[l - 1]));
    const long long tot = (long long)B * c->dl[l - 1] * c->dcp[l - 1] / (16 / c->esz);
    DISPATCH_T(c, glue_launch(ps_scatter_mask_kernel<T>, dim3(grid_for(tot)), dim3(256), 0, c->stream, 
                      (const T*)c->DX[l - 1], (const T*)c->H[l - 1], (T*)c->DA[l - 1], B, group_b, c->dl[l - 1],
                      c->dcp[l - 1], group_shifts(sh, groups, l - 1)));
    CK(post_launch(c, "ps_scatter_mask"));
  }
  CK(pad_out(c, c->DA[l - 1], out, (long long)B * c->dl[l - 1], c->dc[l - 1], c->dcp[l - 1]));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int cg_profile(cg_ctx* c, int enable) {
  c->profiling = enable != 0;
  return 0;
}
extern "C" int cg_profile_report(cg_ctx* c, double out[12]) {
  CU(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 12; ++i) out[i] = 0;
  for (auto& r : c->prof) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, r.e0, r.e1));
    if (getenv("CG_PROF_DUMP")) fprintf(stderr, "[prof] %-60s %8.1f us %7.1f TF/s\n", r.desc, ms * 1e3, r.flops / ms / 1e9);
    out[r.cls * 3 + 0] += ms;
    out[r.cls * 3 + 1] += r.flops;
    out[r.cls * 3 + 2] += 1;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  c->prof.clear();
  return 0;
}

// Per-kernel table of everything timed since cg_profile(ctx, 1): one line per kernel,
//   name <tab> bound (tensor|hbm) <tab> launches <tab> milliseconds <tab> algorithmic FLOPs <tab> algorithmic bytes
// Returns the number of bytes written (truncated to cap - 1), clears the accumulators.
extern "C" int cg_profile_report_text(cg_ctx* c, char* out, int cap) {
  if (!out || cap < 1) return set_err("cg_profile_report_text: bad arguments");
  CU(cudaStreamSynchronize(c->stream));
  struct Agg { int n = 0; double ms = 0, flops = 0, bytes = 0; int cls = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  static const char* cls_name[3] = {"rsgemm_tc", "wgrad_tc", "ghead_tc"};
  for (auto& r : c->prof) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, r.e0, r.e1));
    const std::string key = r.cls < 3 ? cls_name[r.cls] : r.desc;
    if (!agg.count(key)) order.push_back(key);
    Agg& a = agg[key];
    a.n++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes; a.cls = r.cls;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  c->prof.clear();
  std::string txt;
  for (auto& k : order) {
    const Agg& a = agg[k];
    char line[256];
    snprintf(line, sizeof(line), "%s\t%s\t%d\t%.6f\t%.6e\t%.6e\n", k.c_str(), (a.cls == 0 || a.cls == 1) ? "tensor" : "hbm", a.n, a.ms,
             a.flops, a.bytes);
    txt += line;
  }
  const int n = (int)txt.size() < cap - 1 ? (int)txt.size() : cap - 1;
  memcpy(out, txt.data(), n);
  out[n] = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------ microbench hook
extern "C" int cg_bench_layer(cg_ctx* c, int which, int layer, int pass, int B, int iters, float* ms_out,
                              double* flops_out) {
  if (layer < 1 || layer > NL) return set_err("cg_bench_layer: layer must be 1..5");
  if (iters < 1) iters = 1;
  int Bt = B;
  if (which == CG_DISCRIMINATOR) { if (Bt < 1 || Bt > 3 * c->Bmax) return set_err("cg_bench_layer: batch out of range"); }
  else CK(check_batch(c, B));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double macs = 0;
  for (int it = -1; it < iters; ++it) {
    if (it == 0) CU(cudaEventRecord(e0, c->stream));
    if (which == CG_DISCRIMINATOR) {
      const int l = layer;
      macs = (double)Bt * c->dl[l] * c->K * c->dc[l - 1] * c->dc[l];
      if (pass == 0) CK(launch_rsgemm(c, conv_fwd_params(c, l, c->X[l - 1], c->H[l], Bt, EPI_BIAS_LRELU, nullptr)));
      else if (pass == 1) CK(d_dgrad_layer(c, l, 0, l == 1 ? (Bt > c->Bmax ? c->Bmax : Bt) : Bt, l == 1 ? c->DX[0] : c->DX[l - 1]));
      else {
        WgParams w;
        memset(&w, 0, sizeof(w));
        const SegTable st = seg_strided(c->K, c->dcp[l - 1]);
        w.S = c->X[l - 1]; w.s_bs = (long long)c->dl[l - 1] * c->dcp[l - 1]; w.s_rs = 2 * c->dcp[l - 1]; w.s_rows = c->dl[l - 1] / 2;
        w.P = c->DA[l]; w.p_bs = (long long)c->dl[l] * c->dcp[l]; w.p_rs = c->dcp[l];
        w.dW = dgrad(c, 2 * (l - 1)); w.m_real = c->dc[l - 1]; w.n_real = c->dc[l];
        w.B = Bt; w.Q = c->dl[l]; w.Mp = c->dcp[l - 1]; w.Np = c->dcp[l]; w.nseg = c->K;
        for (int k = 0; k < c->K; ++k) { w.shift[k] = st.shift[0][k]; w.scol[k] = st.acol[0][k]; }
        CK(launch_wgrad(c, w));
      }
      if (pass == 1 && l == 1 && Bt > c->Bmax) macs = (double)c->Bmax * c->dl[l] * c->K * c->dc[l - 1] * c->dc[l];
    } else {
      const int i = layer;
      macs = (double)B * c->gl[i - 1] * c->K * c->gc[i - 1] * c->gc[i];
      if (pass != 0) return set_err("cg_bench_layer: generator supports pass 0 only");
      RsParams p;
      memset(&p, 0, sizeof(p));
      p.A = c->HG[i - 1]; p.a_bs = (long long)c->gl[i - 1] * c->gcp[i - 1]; p.a_rs = c->gcp[i - 1]; p.a_rows = c->gl[i - 1];
      p.W = c->Wf_g[i]; p.w_ld = c->K * c->gcp[i - 1];
      p.out = c->AG[i]; p.o_bs = (long long)c->gl[i] * c->gcp[i]; p.o_rs = 2 * c->gcp[i]; p.o_phase_col = c->gcp[i];
      p.bias = gparam(c, c->g_b[i]);
      p.B = B; p.Q = c->gl[i - 1]; p.N = c->gcp[i]; p.n_real = c->gc[i]; p.Kc = c->gcp[i - 1]; p.epi = EPI_BIAS;
      p.seg = seg_transposed(c->K, c->gcp[i - 1]);
      CK(launch_rsgemm(c, p));
    }
  }
  CU(cudaEventRecord(e1, c->stream));
  CU(cudaEventSynchronize(e1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_out) *ms_out = ms / iters;
  if (flops_out) *flops_out = 2.0 * macs;
  return 0;
}
