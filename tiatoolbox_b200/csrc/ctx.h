// ctx.h — per-device handle shared by the subsystems of libtiab200.so
#pragma once
#include <cuda_runtime.h>

#include <string>

struct tiab_ctx {
  int device = 0;
  int sm_count = 148;
  std::string err;
  int stain_path = 0;       // 0 auto, 1 fused, 2 general
  int last_launches = 0;
  // stain workspaces (grown on demand)
  void* stain_work = nullptr;   // StainWork[n]
  size_t stain_work_bytes = 0;
  void* stain_state = nullptr;  // StainState[n]
  size_t stain_state_bytes = 0;
  void* stain_lut = nullptr;    // global constant tables
  // host staging (pinned) + device staging for the *_host entry points
  void* pin_in = nullptr;
  void* pin_out = nullptr;
  size_t pin_bytes = 0;
  void* dev_in = nullptr;
  void* dev_out = nullptr;
  size_t dev_bytes = 0;
  void* dev_small = nullptr;    // N*(6+2) doubles + N ints
  void* pin_small = nullptr;
  size_t small_bytes = 0;
  // HoVer-Net post-processing workspace
  void* pp_work = nullptr;
  size_t pp_bytes = 0;
  int pp_const_set = 0;
  int pp_watershed_impl = 4;  // 1: binary heap, one thread per component (r1c); 2: 8-ary heap, one warp per component; 3: work list + shared memory, lane 0; 4 (default since r2c): work list + shared memory, warp-cooperative 32-ary heap
  int reinhard_path = 0;      // 0: hist + finalise + apply (default); 1: experimental one-CTA-per-image fused kernel when the image fits
  int lab_tables_set = 0;   // csrc/reinhard.cu: __device__ copy of lab_tables.h uploaded
  size_t pp_dbg_off[8] = {0};
  size_t pp_dbg_bytes[8] = {0};
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
};

// Every C entry point works on the handle's device and puts the calling thread's CURRENT device back afterwards: a library that silently
// changes it would make the caller's next allocation (e.g. PyTorch's on another rank's GPU) land on the wrong device.
struct tiab_device_guard {
  int prev = -1;
  bool switched = false;
  explicit tiab_device_guard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    else if (prev < 0) cudaSetDevice(device);
  }
  ~tiab_device_guard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define TIAB_ON_DEVICE(h) tiab_device_guard _tiab_guard((h)->device)

#define TIAB_CUDA_OK(h, expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                           \
      return TIAB_E_CUDA;                                                                      \
    }                                                                                          \
  } while (0)
