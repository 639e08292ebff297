// Error plumbing shared by every translation unit of lib3dhg_sm100a.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace hg {

constexpr int kNumSMsB200 = 148;

void set_error(const char* fmt, ...);  // defined in abi.cu (thread-local message)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

#define HG_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      hg::set_error(__VA_ARGS__);    \
      return 1;                      \
    }                                \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMsB200;
  }
  return n;
}

}  // namespace hg
