// Host helpers of the tensor-core engine: driver entry point for tensor-map encoding (resolved at run time so the
// library does not link against libcuda and loads on a machine without a driver), SM count.
#include "gemm_tc.cuh"

namespace sfno {

std::atomic<int> g_tc_debug{0};

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
    else
      cudaGetLastError();
  }
  return fn;
}

unsigned long long* tc_prof_buffer() {
  static unsigned long long* buf = nullptr;
  if (!buf) {
    if (cudaMalloc(&buf, 12 * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); buf = nullptr; }
    else cudaMemset(buf, 0, 12 * sizeof(unsigned long long));
  }
  return buf;
}

int tc_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace sfno

// measurement hook: read (and clear) the role-wait counters filled by tensor-core launches while tc_debug bit7 is set
extern "C" int sfno_b200_tc_counters(unsigned long long* out12) {
  if (!out12) return SFNO_ERR_INVALID_ARGUMENT;
  unsigned long long* buf = sfno::tc_prof_buffer();
  if (!buf) return SFNO_ERR_CUDA;
  if (cudaDeviceSynchronize() != cudaSuccess) return SFNO_ERR_CUDA;
  if (cudaMemcpy(out12, buf, 12 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return SFNO_ERR_CUDA;
  cudaMemset(buf, 0, 12 * sizeof(unsigned long long));
  return SFNO_OK;
}
