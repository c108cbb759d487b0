// capi.cu — handle management of libtiab200.so
#include <cuda_runtime.h>

#include <new>

#include "../../include/tiab200.h"
#include "ctx.h"

extern "C" {

const char* tiab_version(void) { return "0.1"; }

int tiab_create(int device, tiab_handle_t* out) {
  if (!out) return TIAB_E_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || device < 0 || device >= count) return TIAB_E_CUDA;
  tiab_ctx* h = new (std::nothrow) tiab_ctx();
  if (!h) return TIAB_E_NOMEM;
  h->device = device;
  tiab_device_guard guard(device);   // the caller's current device is restored on return
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != device) {
    delete h;
    return TIAB_E_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return TIAB_E_CUDA;
  }
  *out = h;
  return TIAB_OK;
}

void tiab_destroy(tiab_handle_t h) {
  if (!h) return;
  TIAB_ON_DEVICE(h);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  void* dev[] = {h->stain_work, h->stain_state, h->stain_lut, h->dev_in, h->dev_out, h->dev_small, h->pp_work};
  for (void* p : dev)
    if (p) cudaFree(p);
  void* pin[] = {h->pin_in, h->pin_out, h->pin_small};
  for (void* p : pin)
    if (p) cudaFreeHost(p);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
}

const char* tiab_last_error(tiab_handle_t h) { return h ? h->err.c_str() : "null handle"; }

}  // extern "C"
