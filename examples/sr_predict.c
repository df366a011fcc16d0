/* sr_predict -- model.predict of the DifvdsrDouble graph from plain C, through the graph-level C ABI of libsr100
 * (include/sr100.h: sr_model_create / sr_model_forward).  What a binding in any language does; no Python, no torch.
 *
 * Replaces, for a caller that is not Python: model = DifvdsrDouble(...).create_model(h, w); model.load_weights(...);
 * y = model.predict(x)   (reference models.py:1159-1222, :1217-1218, :342).
 *
 *   sr_predict <params.f32> <input.f32> <NB> <H> <W> <output.f32> [repeat]
 *
 * params.f32: the flat fp32 parameter arena in sr_model_layer order (kernel HWIO then bias per layer, 21,838,211 floats
 * -- tools/export_arena.py writes it from an .h5 / .npz weight file); input.f32: NB*H*W*3 floats in [0,1], NHWC;
 * output.f32: NB*4H*4W*3 floats.  `repeat` runs the forward that many times (the third run on replays a CUDA graph)
 * and prints the mean time of the last ones.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "sr100.h"

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));         \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)
#define SR(call)                                                                          \
  do {                                                                                    \
    int rc_ = (call);                                                                     \
    if (rc_ != SR_OK) {                                                                   \
      fprintf(stderr, "%s:%d: sr100 error %d: %s\n", __FILE__, __LINE__, rc_, sr_last_error_string()); \
      return 3;                                                                           \
    }                                                                                     \
  } while (0)

static float* read_floats(const char* path, size_t n) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    perror(path);
    return NULL;
  }
  float* p = (float*)malloc(n * sizeof(float));
  if (p && fread(p, sizeof(float), n, f) != n) {
    fprintf(stderr, "%s: expected %zu floats\n", path, n);
    free(p);
    p = NULL;
  }
  fclose(f);
  return p;
}

int main(int argc, char** argv) {
  if (argc < 7) {
    fprintf(stderr, "usage: %s params.f32 input.f32 NB H W output.f32 [repeat]\n", argv[0]);
    return 1;
  }
  const int NB = atoi(argv[3]), H = atoi(argv[4]), W = atoi(argv[5]);
  const int repeat = argc > 7 ? atoi(argv[7]) : 1;
  if (NB < 1 || H < 1 || W < 1 || repeat < 1) return 1;
  if (!sr_device_supported()) {
    fprintf(stderr, "no sm_100 device: libsr100 has no CPU path\n");
    return 4;
  }
  const size_t n_params = sr_model_param_count();
  const size_t n_in = (size_t)NB * H * W * 3, n_out = n_in * 16;
  float* h_params = read_floats(argv[1], n_params);
  float* h_in = read_floats(argv[2], n_in);
  if (!h_params || !h_in) return 1;

  float *d_params, *d_in, *d_out;
  CK(cudaMalloc((void**)&d_params, n_params * sizeof(float)));
  CK(cudaMalloc((void**)&d_in, n_in * sizeof(float)));
  CK(cudaMalloc((void**)&d_out, n_out * sizeof(float)));
  CK(cudaMemcpy(d_params, h_params, n_params * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_in, h_in, n_in * sizeof(float), cudaMemcpyHostToDevice));

  sr_model* model = NULL;
  SR(sr_model_create(d_params, NULL, &model));          /* default config: bf16 operands, fp32 LR stream, graphs */

  sr_forward_desc d = {0};
  d.NB = NB, d.H = H, d.W = W;
  d.x = d_in, d.out = d_out;
  d.workspace_bytes = sr_model_forward_workspace_bytes(model, &d);
  CK(cudaMalloc(&d.workspace, d.workspace_bytes));

  cudaStream_t st;
  cudaEvent_t e0, e1;
  CK(cudaStreamCreate(&st));
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int warm = repeat > 3 ? 3 : 0;
  for (int i = 0; i < repeat; ++i) {
    if (i == warm) CK(cudaEventRecord(e0, st));
    SR(sr_model_forward(model, &d, st));
  }
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));

  float* h_out = (float*)malloc(n_out * sizeof(float));
  if (!h_out) return 1;
  CK(cudaMemcpy(h_out, d_out, n_out * sizeof(float), cudaMemcpyDeviceToHost));
  FILE* f = fopen(argv[6], "wb");
  if (!f || fwrite(h_out, sizeof(float), n_out, f) != n_out) {
    perror(argv[6]);
    return 1;
  }
  fclose(f);
  double sum = 0.0;
  for (size_t i = 0; i < n_out; ++i) sum += h_out[i];
  sr_model_run_info info;
  SR(sr_model_forward_info(model, &d, &info));
  printf("{\"NB\": %d, \"H\": %d, \"W\": %d, \"launches\": %d, \"conv_tflop\": %.4f, \"graph_replay\": %d, "
         "\"ms_per_forward\": %.4f, \"output_mean\": %.9g}\n",
         NB, H, W, info.launches, info.conv_flops / 1e12, info.graph_replay, ms / (repeat - warm), sum / (double)n_out);

  sr_model_destroy(model);
  cudaFree(d.workspace);
  cudaFree(d_out);
  cudaFree(d_in);
  cudaFree(d_params);
  free(h_out);
  free(h_in);
  free(h_params);
  return 0;
}
