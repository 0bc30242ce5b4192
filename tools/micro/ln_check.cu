// Standalone check of ln_lrelu_forward_kernel / ln_lrelu_backward_kernel (fp32) against a double-precision host
// implementation:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ln_check ln_check.cu && ./ln_check rows C Cp scale offset
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../calciumgan_b200/csrc/cg_kernels_simt.cuh"

static double rel(const std::vector<double>& a, const std::vector<double>& b) {
  double d = 0, n = 0;
  for (size_t i = 0; i < a.size(); ++i) { d += (a[i] - b[i]) * (a[i] - b[i]); n += b[i] * b[i]; }
  return std::sqrt(d / (n > 0 ? n : 1));
}

int main(int argc, char** argv) {
  const long long rows = argc > 1 ? atoll(argv[1]) : 4096;
  const int C = argc > 2 ? atoi(argv[2]) : 102, Cp = argc > 3 ? atoi(argv[3]) : 128;
  const double scale = argc > 4 ? atof(argv[4]) : 1.0, offset = argc > 5 ? atof(argv[5]) : 0.0;
  std::vector<float> A(rows * Cp, 0.f), DH(rows * Cp, 0.f), gam(C), bet(C);
  srand(1);
  auto rnd = []() { return (double)rand() / RAND_MAX * 2 - 1; };
  for (int c = 0; c < C; ++c) { gam[c] = (float)(1 + 0.3 * rnd()); bet[c] = (float)(0.2 * rnd()); }
  for (long long r = 0; r < rows; ++r)
    for (int c = 0; c < C; ++c) { A[r * Cp + c] = (float)(offset + scale * rnd()); DH[r * Cp + c] = (float)rnd(); }
  float *dA, *dDH, *dH, *dDA, *dmu, *drs, *dg, *db, *dgg, *dgb;
  cudaMalloc(&dA, rows * Cp * 4); cudaMalloc(&dDH, rows * Cp * 4); cudaMalloc(&dH, rows * Cp * 4); cudaMalloc(&dDA, rows * Cp * 4);
  cudaMalloc(&dmu, rows * 4); cudaMalloc(&drs, rows * 4); cudaMalloc(&dg, C * 4); cudaMalloc(&db, C * 4);
  cudaMalloc(&dgg, C * 4); cudaMalloc(&dgb, C * 4);
  cudaMemcpy(dA, A.data(), rows * Cp * 4, cudaMemcpyHostToDevice); cudaMemcpy(dDH, DH.data(), rows * Cp * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dg, gam.data(), C * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, bet.data(), C * 4, cudaMemcpyHostToDevice);
  cudaMemset(dgg, 0, C * 4); cudaMemset(dgb, 0, C * 4);
  const int nvec = Cp / 4;
  const int lpr = nvec > 16 ? 32 : (nvec > 8 ? 16 : 8);
  const int maxv = (nvec + lpr - 1) / lpr;
  long long blocks = (rows * lpr + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (lpr == 32) ln_lrelu_forward_kernel<float, 32><<<(int)blocks, 256>>>(dA, dg, db, dH, dmu, drs, rows, C, Cp);
  else if (lpr == 16) ln_lrelu_forward_kernel<float, 16><<<(int)blocks, 256>>>(dA, dg, db, dH, dmu, drs, rows, C, Cp);
  else ln_lrelu_forward_kernel<float, 8><<<(int)blocks, 256>>>(dA, dg, db, dH, dmu, drs, rows, C, Cp);
  long long bb = (rows * lpr + 255) / 256;
  const long long cap = 148 * (maxv <= 1 ? 6 : 3);
  if (bb > cap) bb = cap;
  const size_t sm = 2 * Cp * sizeof(float);
  if (lpr == 32 && maxv == 1) ln_lrelu_backward_kernel<float, 32, 1><<<(int)bb, 256, sm>>>(dDH, dA, dH, dmu, drs, dg, dDA, dgg, dgb, rows, C, Cp);
  else if (lpr == 32 && maxv == 2) ln_lrelu_backward_kernel<float, 32, 2><<<(int)bb, 256, sm>>>(dDH, dA, dH, dmu, drs, dg, dDA, dgg, dgb, rows, C, Cp);
  else if (lpr == 32) ln_lrelu_backward_kernel<float, 32, 4><<<(int)bb, 256, sm>>>(dDH, dA, dH, dmu, drs, dg, dDA, dgg, dgb, rows, C, Cp);
  else if (lpr == 16) ln_lrelu_backward_kernel<float, 16, 1><<<(int)bb, 256, sm>>>(dDH, dA, dH, dmu, drs, dg, dDA, dgg, dgb, rows, C, Cp);
  else ln_lrelu_backward_kernel<float, 8, 1><<<(int)bb, 256, sm>>>(dDH, dA, dH, dmu, drs, dg, dDA, dgg, dgb, rows, C, Cp);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> H(rows * Cp), DA(rows * Cp), gg(C), gb(C);
  cudaMemcpy(H.data(), dH, rows * Cp * 4, cudaMemcpyDeviceToHost); cudaMemcpy(DA.data(), dDA, rows * Cp * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(gg.data(), dgg, C * 4, cudaMemcpyDeviceToHost); cudaMemcpy(gb.data(), dgb, C * 4, cudaMemcpyDeviceToHost);
  std::vector<double> rH, gH, rDA, gDA, rg(C, 0), rb(C, 0), ggd(gg.begin(), gg.end()), gbd(gb.begin(), gb.end());
  for (long long r = 0; r < rows; ++r) {
    double mu = 0, var = 0;
    for (int c = 0; c < C; ++c) mu += A[r * Cp + c];
    mu /= C;
    for (int c = 0; c < C; ++c) { const double d = A[r * Cp + c] - mu; var += d * d; }
    const double rstd = 1.0 / std::sqrt(var / C + 1e-3);
    std::vector<double> xh(C), dn(C);
    double s1 = 0, s2 = 0;
    for (int c = 0; c < C; ++c) {
      xh[c] = (A[r * Cp + c] - mu) * rstd;
      const double y = gam[c] * xh[c] + bet[c];
      const double h = y > 0 ? y : 0.3 * y;
      rH.push_back(h); gH.push_back(H[r * Cp + c]);
      dn[c] = DH[r * Cp + c] * (h > 0 ? 1.0 : 0.3);
      s1 += gam[c] * dn[c]; s2 += gam[c] * dn[c] * xh[c];
      rg[c] += dn[c] * xh[c]; rb[c] += dn[c];
    }
    s1 /= C; s2 /= C;
    for (int c = 0; c < C; ++c) { rDA.push_back(rstd * (gam[c] * dn[c] - s1 - xh[c] * s2)); gDA.push_back(DA[r * Cp + c]); }
  }
  printf("rows %lld C %d Cp %d scale %g offset %g | H %.2e DA %.2e dgamma %.2e dbeta %.2e\n", rows, C, Cp, scale, offset,
         rel(gH, rH), rel(gDA, rDA), rel(ggd, rg), rel(gbd, rb));
  return 0;
}
