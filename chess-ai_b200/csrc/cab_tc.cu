// cab_tc.cu — host side of the bf16 tcgen05 path (cab_eval_bf16_*): weight packing, TMA tensor maps, layer loop.
#include <cuda.h>

#include <cstdio>

#include "cab_internal.cuh"
#include "tc_gemm_bf16.cuh"

namespace cab {

struct TcState {
  std::vector<__nv_bfloat16 *> wt;  // per GEMM layer: Wt[N][K]
  float *w_last = nullptr;          // last layer (N = 1) in fp32
  __nv_bfloat16 *act[2] = {nullptr, nullptr};
  double *d_out = nullptr;
  int8_t *d_codes = nullptr;
  long cap = 0;
  int max_width = 0;
  bool packed = false;
  cudaStream_t stream = nullptr;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return CAB_OK;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr) { set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e)); return CAB_ECUDA; }
  g_encode = (EncodeTiledFn) fn;
  return CAB_OK;
}

// row-major bf16 matrix [rows][cols] (cols contiguous) -> tiled map with box {64 cols, box_rows}, 128-byte swizzle
static int make_map(CUtensorMap *m, const void *ptr, long rows, long cols, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t) cols, (cuuint64_t) rows};
  const cuuint64_t gstride[1] = {(cuuint64_t) cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t) tc::BK, (cuuint32_t) box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%ld cols=%ld", (int) r, rows, cols); return CAB_ECUDA; }
  return CAB_OK;
}

static bool tc_supported(const cab_net *net) {
  const int L = (int) net->dims.size();
  if (L < 3 || net->dims[0] != 64 || net->dims[L - 1] != 1) return false;
  for (int l = 1; l + 1 < L; ++l)
    if (net->dims[l] % tc::BN != 0) return false;
  return true;
}

void tc_free(cab_net *net) {
  TcState *s = static_cast<TcState *>(net->tc);
  if (!s) return;
  for (auto *p : s->wt) cudaFree(p);
  cudaFree(s->w_last); cudaFree(s->act[0]); cudaFree(s->act[1]); cudaFree(s->d_out); cudaFree(s->d_codes);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  net->tc = nullptr;
}

static int tc_prepare(cab_net *net, long n, cudaStream_t stream) {
  if (!tc_supported(net)) {
    set_error("bf16 tcgen05 path needs hidden widths that are multiples of %d (got a net it does not fit); use the fp64 path", tc::BN);
    return CAB_EINVAL;
  }
  int rc = load_encode();
  if (rc != CAB_OK) return rc;
  TcState *s = static_cast<TcState *>(net->tc);
  const int L = (int) net->dims.size();
  if (!s) {
    s = new TcState();
    net->tc = s;
    for (int l = 0; l + 2 < L; ++l) {
      __nv_bfloat16 *p = nullptr;
      CAB_CUDA(cudaMalloc(&p, (size_t) net->dims[l] * net->dims[l + 1] * 2));
      s->wt.push_back(p);
      s->max_width = std::max(s->max_width, net->dims[l + 1]);
    }
    CAB_CUDA(cudaMalloc(&s->w_last, (size_t) net->dims[L - 2] * sizeof(float)));
    CAB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CAB_CUDA(cudaFuncSetAttribute(tc::tc_gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tc::kSmemTc));
  }
  if (stream == nullptr) stream = s->stream;  // host entry point: everything on the state's own stream
  if (!s->packed || net->tc_dirty) {
    for (int l = 0; l + 2 < L; ++l) {
      const long total = (long) net->dims[l] * net->dims[l + 1];
      tc::pack_wt_bf16_kernel<<<(int) std::min<long>((total + 255) / 256, 148L * 16), 256, 0, stream>>>(
          net->d_w + net->woff[l], net->dims[l], net->dims[l + 1], s->wt[l]);
    }
    tc::pack_last_f32_kernel<<<8, 256, 0, stream>>>(net->d_w + net->woff[L - 2], net->dims[L - 2], s->w_last);
    net->launches += L - 1;
    CAB_CUDA(cudaGetLastError());
    CAB_CUDA(cudaStreamSynchronize(stream));  // rare (weights changed): later launches may sit on another stream
    s->packed = true;
    net->tc_dirty = false;
  }
  if (n > s->cap) {
    cudaFree(s->act[0]); cudaFree(s->act[1]); cudaFree(s->d_out); cudaFree(s->d_codes);
    const long pad = (n + tc::BM - 1) / tc::BM * tc::BM;
    CAB_CUDA(cudaMalloc(&s->act[0], (size_t) pad * std::max(s->max_width, 64) * 2));
    CAB_CUDA(cudaMalloc(&s->act[1], (size_t) pad * std::max(s->max_width, 64) * 2));
    CAB_CUDA(cudaMalloc(&s->d_out, pad * sizeof(double)));
    CAB_CUDA(cudaMalloc(&s->d_codes, pad * 64));
    s->cap = pad;
  }
  return CAB_OK;
}

static int tc_forward(cab_net *net, const int8_t *d_codes, long n, double *d_out, cudaStream_t stream) {
  int rc = tc_prepare(net, n, stream);
  if (rc != CAB_OK) return rc;
  TcState *s = static_cast<TcState *>(net->tc);
  const int L = (int) net->dims.size();
  tc::codes_to_bf16_kernel<<<(int) std::min<long>((n * 64 + 255) / 256, 148L * 16), 256, 0, stream>>>(d_codes, n * 64, s->act[0]);
  net->launches++;
  int cur = 0;
  for (int l = 0; l + 2 < L; ++l) {
    const int K = net->dims[l], N = net->dims[l + 1];
    CUtensorMap mx, mw;
    if ((rc = make_map(&mx, s->act[cur], n, K, tc::BM)) != CAB_OK) return rc;
    if ((rc = make_map(&mw, s->wt[l], N, K, tc::BN)) != CAB_OK) return rc;
    tc::TcGemmArgs a{(int) n, N, K, s->act[1 - cur], 1};
    const int tiles = (int) ((n + tc::BM - 1) / tc::BM) * (N / tc::BN);
    tc::tc_gemm_bf16_kernel<<<std::min(tiles, net->sm_count), tc::kThreadsTc, tc::kSmemTc, stream>>>(mx, mw, a);
    net->launches++;
    CAB_CUDA(cudaGetLastError());
    cur = 1 - cur;
  }
  const int K = net->dims[L - 2];
  tc::gemv_out_kernel<<<(int) ((n * 32 + 255) / 256), 256, 0, stream>>>(s->act[cur], s->w_last, n, K, d_out);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  return CAB_OK;
}

}  // namespace cab

using namespace cab;

extern "C" {

int cab_eval_bf16_dev(cab_net *net, const int8_t *codes_dev, long n, double *out_dev, void *stream) {
  if (!net || n < 0 || (n > 0 && (!codes_dev || !out_dev))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  CAB_CUDA(cudaSetDevice(net->device));
  return tc_forward(net, codes_dev, n, out_dev, (cudaStream_t) stream);
}

int cab_eval_bf16_host(cab_net *net, const int8_t *codes, long n, double *out) {
  if (!net || n < 0 || (n > 0 && (!codes || !out))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = tc_prepare(net, n, nullptr);
  if (rc != CAB_OK) return rc;
  TcState *s = static_cast<TcState *>(net->tc);
  CAB_CUDA(cudaMemcpyAsync(s->d_codes, codes, n * 64, cudaMemcpyHostToDevice, s->stream));
  rc = tc_forward(net, s->d_codes, n, s->d_out, s->stream);
  if (rc != CAB_OK) return rc;
  CAB_CUDA(cudaMemcpyAsync(out, s->d_out, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CAB_CUDA(cudaStreamSynchronize(s->stream));
  return CAB_OK;
}

}  // extern "C"
