// tcgen05 / TMEM / TMA implicit-GEMM kernels (bf16 mixed-precision path). Placeholder until the
// kernels land: reports "unsupported" so the engine uses the CUDA-core kernels.
#pragma once
#include "cg_common.cuh"

struct TcState { int dummy; };
static inline int tc_init(TcState*) { return 0; }
static inline void tc_destroy(TcState*) {}
static inline bool tc_rsgemm_supported(const RsParams&) { return false; }
static inline int tc_rsgemm_launch(TcState*, const RsParams&, cudaStream_t) { return 1; }
static inline bool tc_wgrad_supported(const WgParams&) { return false; }
static inline int tc_wgrad_launch(TcState*, const WgParams&, cudaStream_t) { return 1; }
