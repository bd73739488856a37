// The handful of CUDA runtime entry points csrc/flamingo_b200.cu calls, for the host emulation build
// (tests/cpu_harness): "device memory" is host memory, streams and events do nothing (every emulated launch is
// synchronous), the device reports compute capability 10.x with FM_EMU_SMS multiprocessors (default 4, so persistent
// kernels loop over several units per CTA), and cuTensorMapEncodeTiled resolves to the emulator's encoder.
// TEST INFRASTRUCTURE ONLY — never linked into the shipped library.
#include <cstdlib>
#include <cstring>

#include "tc_emu.h"

extern "C" {

cudaError_t cudaGetDevice(int* dev) { *dev = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* value, enum cudaDeviceAttr attr, int) {
  if (attr == cudaDevAttrComputeCapabilityMajor) { *value = 10; return cudaSuccess; }
  if (attr == cudaDevAttrMultiProcessorCount) {
    const char* e = std::getenv("FM_EMU_SMS");
    *value = e ? std::atoi(e) : 4;
    if (*value < 1) *value = 1;
    return cudaSuccess;
  }
  return cudaErrorInvalidValue;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaFuncSetAttribute(const void*, enum cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int) {
  static int side_stream_token;
  *s = reinterpret_cast<cudaStream_t>(&side_stream_token);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpyFromSymbol(void* dst, const void* symbol, size_t n, size_t off, enum cudaMemcpyKind) {
  std::memcpy(dst, static_cast<const char*>(symbol) + off, n);
  return cudaSuccess;
}
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, enum cudaDriverEntryPointQueryResult* st) {
  if (std::strcmp(symbol, "cuTensorMapEncodeTiled") != 0) { if (st) *st = cudaDriverEntryPointSymbolNotFound; *fn = nullptr; return cudaSuccess; }
  *fn = reinterpret_cast<void*>(&emu::encode_tiled);
  if (st) *st = cudaDriverEntryPointSuccess;
  return cudaSuccess;
}

}  // extern "C"
