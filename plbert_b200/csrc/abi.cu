// plbert_b200 — library plumbing behind the C ABI: thread-local error slot, launch checks,
// the TMA descriptor cache, and the CUDA-core check GEMM used by the tests.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace plb {

namespace {
thread_local char g_error[512] = "";

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime so the library carries no link-time dependency on libcuda
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct MapKey {
  uintptr_t base;
  uint64_t rows, cols, ld;
  uint32_t box_rows, box_cols;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           box_cols == o.box_cols;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = k.base * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.ld + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= ((uint64_t(k.box_rows) << 32 | k.box_cols) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    return static_cast<size_t>(h);
  }
};
std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return PLB_ERR_CUDA;
  }
  return PLB_OK;
}

static int g_launch_overlap = 1;
bool launch_overlap_enabled() { return g_launch_overlap != 0; }
void set_launch_overlap(int on) { g_launch_overlap = on ? 1 : 0; }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

bool tensor_map_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows, uint32_t box_cols) {
  MapKey key{reinterpret_cast<uintptr_t>(base), rows, cols, ld, box_rows, box_cols};
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return false;
  }
  CUtensorMap m;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {ld * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for base=%p rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r, base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return false;
  }
  if (g_maps.size() > 4096) g_maps.clear();  // maps are returned by value, so dropping the cache is safe
  g_maps.emplace(key, m);
  *out = m;
  return true;
}

// ------------------------------------------------------------------------------------------------
// check GEMM: one thread per output element, fp32 math.  Test infrastructure only.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void gemm_check_kernel(int op, int epi, int M, int N, int K, const __nv_bfloat16* A, int lda,
                                  const __nv_bfloat16* B, int ldb, void* C, int ldc, const float* bias,
                                  const __nv_bfloat16* aux_in, int ld_in, __nv_bfloat16* aux_out, int ld_out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= M || n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float a, b;
    if (op == PLB_GEMM_FWD) {
      a = __bfloat162float(A[(size_t)m * lda + k]);
      b = __bfloat162float(B[(size_t)n * ldb + k]);
    } else if (op == PLB_GEMM_DGRAD) {
      a = __bfloat162float(A[(size_t)m * lda + k]);
      b = __bfloat162float(B[(size_t)k * ldb + n]);
    } else {
      a = __bfloat162float(A[(size_t)k * lda + m]);
      b = __bfloat162float(B[(size_t)k * ldb + n]);
    }
    acc = fmaf(a, b, acc);
  }
  if (op == PLB_GEMM_WGRAD) {
    static_cast<float*>(C)[(size_t)m * ldc + n] += acc;
    return;
  }
  if (bias) acc += bias[n];
  __nv_bfloat16* c = static_cast<__nv_bfloat16*>(C);
  if (epi == PLB_EPI_RESIDUAL) acc += __bfloat162float(aux_in[(size_t)m * ld_in + n]);
  if (epi == PLB_EPI_GELU_BWD) acc *= __bfloat162float(aux_in[(size_t)m * ld_in + n]);
  if (epi == PLB_EPI_GELU) {
    const float u = __bfloat162float(__float2bfloat16(acc));
    aux_out[(size_t)m * ld_out + n] = __float2bfloat16(gelu_new_grad(u));
    acc = gelu_new_fwd(u);
  }
  c[(size_t)m * ldc + n] = __float2bfloat16(acc);
}

__global__ void length_to_mask_kernel(const int32_t* lengths, int B, int L, int32_t* mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const int b = i / L, j = i - b * L;
  mask[i] = ((j + 1) > lengths[b]) ? 0 : 1;
}
}  // namespace

}  // namespace plb

using namespace plb;

extern "C" int plb_abi_version(void) { return PLB_ABI_VERSION; }
extern "C" const char* plb_last_error_string(void) { return g_error; }
namespace plb {
namespace {
// p[0 .. bytes) = 0: 16-byte stores over the aligned middle, byte stores at the ragged ends
__global__ void __launch_bounds__(256) fill_zero_kernel(uint8_t* __restrict__ p, size_t bytes) {
  pdl_grid_sync();
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
  const size_t h = head < bytes ? head : bytes;
  const size_t nvec = (bytes - h) >> 4;
  uint4* v = reinterpret_cast<uint4*>(p + h);
  const size_t tid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x, nthr = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = tid; i < nvec; i += nthr) v[i] = make_uint4(0, 0, 0, 0);
  const size_t tail0 = h + (nvec << 4);
  if (tid < h) p[tid] = 0;
  if (tid < bytes - tail0) p[tail0 + tid] = 0;
}
}  // namespace
}  // namespace plb

extern "C" int plb_fill_zero(void* dst, unsigned long long bytes, plb_stream_t stream) {
  PLB_REQUIRE(dst != nullptr || bytes == 0, PLB_ERR_ARG, "fill_zero: null pointer");
  if (bytes == 0) return PLB_OK;
  const unsigned long long want = (bytes / 16 + 255) / 256 + 1;
  const int grid = static_cast<int>(want < plb::sm_count() * 8ull ? want : plb::sm_count() * 8ull);
  cudaError_t e = plb::launch_kernel(plb::fill_zero_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream),
                                     static_cast<uint8_t*>(dst), static_cast<size_t>(bytes));
  PLB_REQUIRE(e == cudaSuccess, PLB_ERR_CUDA, "plb_fill_zero: launch failed: %s", cudaGetErrorString(e));
  return plb::check_launch("plb_fill_zero");
}

extern "C" int plb_set_launch_overlap(int enable) {
  plb::set_launch_overlap(enable);
  return PLB_OK;
}

extern "C" int plb_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    set_error("no CUDA device");
    return PLB_ERR_CUDA;
  }
  return n;
}

extern "C" int plb_length_to_mask(const int32_t* lengths, int B, int L, int32_t* attn_mask, plb_stream_t stream) {
  PLB_REQUIRE(lengths && attn_mask && B > 0 && L > 0, PLB_ERR_ARG, "length_to_mask: bad arguments");
  const int n = B * L;
  length_to_mask_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(lengths, B, L, attn_mask);
  return check_launch("plb_length_to_mask");
}

extern "C" int plb_gemm_bf16_check(int op, int epilogue, int M, int N, int K, const void* A, int lda, const void* B,
                                   int ldb, void* C, int ldc, const float* bias, const void* aux_in, int ld_aux_in,
                                   void* aux_out, int ld_aux_out, plb_stream_t stream) {
  PLB_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && C, PLB_ERR_ARG, "gemm_check: bad arguments");
  dim3 block(32, 8), grid((N + 31) / 32, (M + 7) / 8);
  gemm_check_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      op, epilogue, M, N, K, static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(B), ldb, C,
      ldc, bias, static_cast<const __nv_bfloat16*>(aux_in), ld_aux_in, static_cast<__nv_bfloat16*>(aux_out),
      ld_aux_out);
  return check_launch("plb_gemm_bf16_check");
}
