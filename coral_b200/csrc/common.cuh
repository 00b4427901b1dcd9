// Shared host helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/coral_b200.h"

namespace coral {

void set_error(const std::string& msg);
int32_t fail(int32_t code, const std::string& msg);

#define CORAL_CUDA_OK(expr)                                                                 \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::coral::fail(CORAL_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (dev >= 0 && dev != prev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

int sm_count(int device);

}  // namespace coral
