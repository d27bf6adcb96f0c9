// Standalone device self-test for the tcgen05 GEMM (no Python, no torch): every (op, epilogue)
// pair the library exports is compared against the CUDA-core check kernel on ragged and aligned
// shapes, then the training shapes are timed.  Built by plbert_b200/csrc/build.py; run on a B200:
//   tests/native/gemm_selftest [--perf]
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/plbert_b200.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static uint32_t g_seed = 12345;
static float frand() {
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFF) / 65536.0f * 2.0f - 1.0f;
}

struct Bf16Mat {
  __nv_bfloat16* d = nullptr;
  int rows, cols, ld;
  Bf16Mat(int r, int c, int ld_, bool random) : rows(r), cols(c), ld(ld_) {
    std::vector<__nv_bfloat16> h((size_t)r * ld_);
    for (auto& v : h) v = __float2bfloat16(random ? frand() : 0.f);
    CK(cudaMalloc(&d, h.size() * 2));
    CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  }
  ~Bf16Mat() { cudaFree(d); }
  std::vector<float> host() const {
    std::vector<__nv_bfloat16> h((size_t)rows * ld);
    CK(cudaMemcpy(h.data(), d, h.size() * 2, cudaMemcpyDeviceToHost));
    std::vector<float> f(h.size());
    for (size_t i = 0; i < h.size(); ++i) f[i] = __bfloat162float(h[i]);
    return f;
  }
};

static int g_fail = 0;

static void compare(const char* name, const std::vector<float>& a, const std::vector<float>& b, int rows, int cols,
                    int ld, float atol, float rtol) {
  double max_err = 0, max_ref = 0;
  long bad = 0;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const float x = a[(size_t)r * ld + c], y = b[(size_t)r * ld + c];
      const float err = fabsf(x - y);
      if (!(err <= atol + rtol * fabsf(y))) {
        if (bad < 5) printf("    mismatch [%d,%d]: got %g want %g\n", r, c, x, y);
        ++bad;
      }
      if (err > max_err) max_err = err;
      if (fabsf(y) > max_ref) max_ref = fabsf(y);
    }
  printf("%-44s max|err| %.4g  max|ref| %.4g  bad %ld  %s\n", name, max_err, max_ref, bad, bad ? "FAIL" : "ok");
  if (bad) ++g_fail;
}

static void run_case(int op, int epi, int M, int N, int K, bool with_bias) {
  char name[128];
  snprintf(name, sizeof(name), "op%d epi%d M%d N%d K%d bias%d", op, epi, M, N, K, (int)with_bias);
  auto pad8 = [](int x) { return (x + 7) / 8 * 8; };
  int a_rows, a_cols, b_rows, b_cols;
  if (op == PLB_GEMM_FWD) { a_rows = M; a_cols = K; b_rows = N; b_cols = K; }
  else if (op == PLB_GEMM_DGRAD) { a_rows = M; a_cols = K; b_rows = K; b_cols = N; }
  else { a_rows = K; a_cols = M; b_rows = K; b_cols = N; }
  Bf16Mat A(a_rows, a_cols, pad8(a_cols), true), B(b_rows, b_cols, pad8(b_cols), true);
  const int ldn = pad8(N);
  std::vector<float> hb(N);
  for (auto& v : hb) v = frand();
  float* bias = nullptr;
  if (with_bias) {
    CK(cudaMalloc(&bias, N * 4));
    CK(cudaMemcpy(bias, hb.data(), N * 4, cudaMemcpyHostToDevice));
  }
  if (op == PLB_GEMM_WGRAD) {
    float *c1, *c2;
    std::vector<float> init((size_t)M * ldn);
    for (auto& v : init) v = frand();
    CK(cudaMalloc(&c1, init.size() * 4));
    CK(cudaMalloc(&c2, init.size() * 4));
    CK(cudaMemcpy(c1, init.data(), init.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c2, init.data(), init.size() * 4, cudaMemcpyHostToDevice));
    int r1 = plb_gemm_bf16(op, epi, M, N, K, A.d, A.ld, B.d, B.ld, c1, ldn, nullptr, nullptr, 0, nullptr, 0, nullptr);
    int r2 = plb_gemm_bf16_check(op, epi, M, N, K, A.d, A.ld, B.d, B.ld, c2, ldn, nullptr, nullptr, 0, nullptr, 0, nullptr);
    if (r1 || r2) { printf("%s: launch error %d %d: %s\n", name, r1, r2, plb_last_error_string()); ++g_fail; return; }
    CK(cudaDeviceSynchronize());
    std::vector<float> h1(init.size()), h2(init.size());
    CK(cudaMemcpy(h1.data(), c1, h1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), c2, h2.size() * 4, cudaMemcpyDeviceToHost));
    compare(name, h1, h2, M, N, ldn, 1e-2f * sqrtf((float)K) * 0.05f + 1e-3f, 1e-3f);
    cudaFree(c1); cudaFree(c2);
    return;
  }
  Bf16Mat C1(M, N, ldn, false), C2(M, N, ldn, false), AuxIn(M, N, ldn, true), U1(M, N, ldn, false), U2(M, N, ldn, false);
  const void* aux_in = (epi == PLB_EPI_RESIDUAL || epi == PLB_EPI_GELU_BWD) ? AuxIn.d : nullptr;
  int r1 = plb_gemm_bf16(op, epi, M, N, K, A.d, A.ld, B.d, B.ld, C1.d, ldn, bias, aux_in, ldn,
                         epi == PLB_EPI_GELU ? U1.d : nullptr, ldn, nullptr);
  int r2 = plb_gemm_bf16_check(op, epi, M, N, K, A.d, A.ld, B.d, B.ld, C2.d, ldn, bias, aux_in, ldn,
                               epi == PLB_EPI_GELU ? U2.d : nullptr, ldn, nullptr);
  if (r1 || r2) { printf("%s: launch error %d %d: %s\n", name, r1, r2, plb_last_error_string()); ++g_fail; return; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: execution failed: %s\n", name, cudaGetErrorString(e)); exit(3); }
  const float atol = 0.02f * sqrtf((float)K) / 3.f + 0.02f;
  compare(name, C1.host(), C2.host(), M, N, ldn, atol, 0.02f);
  if (epi == PLB_EPI_GELU) compare("    (saved gelu derivative)", U1.host(), U2.host(), M, N, ldn, atol, 0.02f);
  if (bias) cudaFree(bias);
}

static const char* g_only = nullptr;  // --only <substring>: time just the matching perf cases (and skip the checks)

static void perf_case(const char* label, int op, int epi, int M, int N, int K) {
  if (g_only && !strstr(label, g_only)) return;
  int a_rows, a_cols, b_rows, b_cols;
  if (op == PLB_GEMM_FWD) { a_rows = M; a_cols = K; b_rows = N; b_cols = K; }
  else if (op == PLB_GEMM_DGRAD) { a_rows = M; a_cols = K; b_rows = K; b_cols = N; }
  else { a_rows = K; a_cols = M; b_rows = K; b_cols = N; }
  Bf16Mat A(a_rows, a_cols, a_cols, true), B(b_rows, b_cols, b_cols, true);
  void* C;
  CK(cudaMalloc(&C, (size_t)M * N * 4));
  CK(cudaMemset(C, 0, (size_t)M * N * 4));
  Bf16Mat* aux = nullptr;
  Bf16Mat* uout = nullptr;
  if (epi == PLB_EPI_RESIDUAL || epi == PLB_EPI_GELU_BWD) aux = new Bf16Mat(M, N, N, true);
  if (epi == PLB_EPI_GELU) uout = new Bf16Mat(M, N, N, false);
  float* bias = nullptr;
  if (op == PLB_GEMM_FWD) { CK(cudaMalloc(&bias, N * 4)); CK(cudaMemset(bias, 0, N * 4)); }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int iters = 10;
  for (int it = 0; it < 3 + iters; ++it) {
    if (it == 3) CK(cudaEventRecord(e0));
    int r = plb_gemm_bf16(op, epi, M, N, K, A.d, A.ld, B.d, B.ld, C, N, bias, aux ? aux->d : nullptr, N,
                          uout ? uout->d : nullptr, N, nullptr);
    if (r) { printf("%s: error %s\n", label, plb_last_error_string()); ++g_fail; return; }
  }
  CK(cudaEventRecord(e1));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: execution failed: %s\n", label, cudaGetErrorString(e)); exit(3); }
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  printf("%-36s M%-6d N%-5d K%-6d %8.3f ms  %8.1f TFLOP/s\n", label, M, N, K, ms, 2.0 * M * N * K / ms * 1e-9);
  cudaFree(C);
  if (bias) cudaFree(bias);
  delete aux;
  delete uout;
}

int main(int argc, char** argv) {
  bool perf = argc > 1 && !strcmp(argv[1], "--perf");
  if (perf && argc > 3 && !strcmp(argv[2], "--only")) g_only = argv[3];
  printf("plbert_b200 ABI %d, SMs %d\n", plb_abi_version(), plb_device_sm_count());
  if (!g_only) {
  // aligned single tile, then multi-tile / multi-unit, then ragged shapes (TMA zero fill + clipping)
  run_case(PLB_GEMM_FWD, PLB_EPI_NONE, 128, 256, 64, false);
  run_case(PLB_GEMM_FWD, PLB_EPI_NONE, 128, 256, 256, true);
  run_case(PLB_GEMM_FWD, PLB_EPI_NONE, 1024, 768, 768, true);
  run_case(PLB_GEMM_FWD, PLB_EPI_NONE, 1000, 188, 768, true);
  run_case(PLB_GEMM_FWD, PLB_EPI_NONE, 40000, 512, 128, true);
  run_case(PLB_GEMM_FWD, PLB_EPI_RESIDUAL, 1024, 768, 768, true);
  run_case(PLB_GEMM_FWD, PLB_EPI_RESIDUAL, 30000, 768, 192, true);
  run_case(PLB_GEMM_FWD, PLB_EPI_GELU, 1024, 2048, 768, true);
  run_case(PLB_GEMM_FWD, PLB_EPI_GELU, 20000, 520, 136, true);
  run_case(PLB_GEMM_DGRAD, PLB_EPI_NONE, 128, 256, 64, false);
  run_case(PLB_GEMM_DGRAD, PLB_EPI_NONE, 1024, 768, 2304, false);
  run_case(PLB_GEMM_DGRAD, PLB_EPI_NONE, 999, 768, 188, false);
  run_case(PLB_GEMM_DGRAD, PLB_EPI_RESIDUAL, 2048, 768, 2048, false);
  run_case(PLB_GEMM_DGRAD, PLB_EPI_GELU_BWD, 2048, 2048, 768, false);
  run_case(PLB_GEMM_DGRAD, PLB_EPI_NONE, 3000, 128, 768, false);
  run_case(PLB_GEMM_WGRAD, PLB_EPI_NONE, 128, 256, 64, false);
  run_case(PLB_GEMM_WGRAD, PLB_EPI_NONE, 768, 768, 4096, false);
  run_case(PLB_GEMM_WGRAD, PLB_EPI_NONE, 188, 768, 1000, false);
  run_case(PLB_GEMM_WGRAD, PLB_EPI_NONE, 2304, 768, 20000, false);
  run_case(PLB_GEMM_WGRAD, PLB_EPI_NONE, 768, 128, 5000, false);
  printf("%s (%d failing cases)\n", g_fail ? "SELFTEST FAILED" : "SELFTEST PASSED", g_fail);
  }
  if (perf) {
    const int T = 98304;
    perf_case("fwd qkv   (bias)", PLB_GEMM_FWD, PLB_EPI_NONE, T, 2304, 768);
    perf_case("fwd o     (bias+residual)", PLB_GEMM_FWD, PLB_EPI_RESIDUAL, T, 768, 768);
    perf_case("fwd ffn1  (bias+gelu)", PLB_GEMM_FWD, PLB_EPI_GELU, T, 2048, 768);
    perf_case("fwd ffn2  (bias+residual)", PLB_GEMM_FWD, PLB_EPI_RESIDUAL, T, 768, 2048);
    perf_case("fwd map   (bias)", PLB_GEMM_FWD, PLB_EPI_NONE, T, 768, 128);
    perf_case("dgrad ffn2 (gelu')", PLB_GEMM_DGRAD, PLB_EPI_GELU_BWD, T, 2048, 768);
    perf_case("dgrad ffn1 (+residual)", PLB_GEMM_DGRAD, PLB_EPI_RESIDUAL, T, 768, 2048);
    perf_case("dgrad o", PLB_GEMM_DGRAD, PLB_EPI_NONE, T, 768, 768);
    perf_case("dgrad qkv (+residual)", PLB_GEMM_DGRAD, PLB_EPI_RESIDUAL, T, 768, 2304);
    perf_case("wgrad qkv", PLB_GEMM_WGRAD, PLB_EPI_NONE, 2304, 768, T);
    perf_case("wgrad o", PLB_GEMM_WGRAD, PLB_EPI_NONE, 768, 768, T);
    perf_case("wgrad ffn1", PLB_GEMM_WGRAD, PLB_EPI_NONE, 2048, 768, T);
    perf_case("wgrad ffn2", PLB_GEMM_WGRAD, PLB_EPI_NONE, 768, 2048, T);
    // the 64 k token head on the masked rows of one batch (M = 11421 rows)
    const int Mr = 11421, Vt = 64000;
    perf_case("fwd token head (bias)", PLB_GEMM_FWD, PLB_EPI_NONE, Mr, Vt, 768);
    perf_case("dgrad token head", PLB_GEMM_DGRAD, PLB_EPI_NONE, Mr, 768, Vt);
    perf_case("wgrad token head", PLB_GEMM_WGRAD, PLB_EPI_NONE, Vt, 768, Mr);
  }
  return g_fail ? 1 : 0;
}
