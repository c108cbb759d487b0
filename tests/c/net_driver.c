/* net_driver.c — run a tiab200 plan file from C: no Python, no torch.
 *   net_driver <plan.bin> <tiles.u8> <out-prefix>
 * reads N = batch uint8 NHWC tiles, runs tiab_net_infer and writes every output head as <out-prefix><k>.f32.
 * This is the binding a cgo / JNI / Rust host would write against include/tiab200.h. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "tiab200.h"

static void* slurp(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  *n = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  void* p = malloc(*n ? *n : 1);
  if (p && fread(p, 1, *n, f) != *n) {
    free(p);
    p = NULL;
  }
  fclose(f);
  return p;
}

int main(int argc, char** argv) {
  if (argc != 4) return 2;
  size_t plan_bytes = 0, tile_bytes = 0;
  void* plan = slurp(argv[1], &plan_bytes);
  void* tiles = slurp(argv[2], &tile_bytes);
  if (!plan || !tiles) return 3;
  tiab_handle_t h = NULL;
  if (tiab_create(0, &h) != 0) return 4;
  tiab_net_t net = NULL;
  if (tiab_net_create(h, plan, plan_bytes, &net) != 0) {
    fprintf(stderr, "%s\n", tiab_last_error(h));
    return 5;
  }
  int info[64];
  if (tiab_net_info(net, info, 64) != 0) return 6;
  const int batch = info[0], n_out = info[4];
  const size_t in_bytes = (size_t)batch * info[1] * info[2] * info[3];
  if (tile_bytes != in_bytes) return 7;
  void* d_in = NULL;
  void* d_out[8] = {0};
  size_t out_bytes[8];
  if (cudaMalloc(&d_in, in_bytes) != cudaSuccess) return 8;
  cudaMemcpy(d_in, tiles, in_bytes, cudaMemcpyHostToDevice);
  for (int k = 0; k < n_out; ++k) {
    size_t n = 4;
    for (int d = 0; d < info[5 + 5 * k]; ++d) n *= (size_t)info[5 + 5 * k + 1 + d];
    out_bytes[k] = n;
    if (cudaMalloc(&d_out[k], n) != cudaSuccess) return 9;
  }
  if (tiab_net_infer(net, (const uint8_t*)d_in, batch, d_out, NULL) != 0) {
    fprintf(stderr, "%s\n", tiab_last_error(h));
    return 10;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return 11;
  for (int k = 0; k < n_out; ++k) {
    void* host = malloc(out_bytes[k]);
    cudaMemcpy(host, d_out[k], out_bytes[k], cudaMemcpyDeviceToHost);
    char path[1024];
    snprintf(path, sizeof path, "%s%d.f32", argv[3], k);
    FILE* f = fopen(path, "wb");
    if (!f || fwrite(host, 1, out_bytes[k], f) != out_bytes[k]) return 12;
    fclose(f);
    free(host);
  }
  tiab_net_destroy(net);
  tiab_destroy(h);
  printf("ok: %d tiles, %d heads\n", batch, n_out);
  return 0;
}
