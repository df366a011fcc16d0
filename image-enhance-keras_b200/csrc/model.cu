// Graph-level entry points of the sr100 C ABI: the whole DifvdsrDouble launch sequence (forward, backward, Adam)
// behind sr_model_forward / sr_model_forward_backward / sr_model_train_step.
//
// Reference: the graph of DifvdsrDouble.create_model (models.py:1159-1222; blocks :1231-1270, scalar lambdas
// :977-986, bilinear :1392-1399), run by model.predict (models.py:342) and trained by fit_generator
// (models.py:146-157) after compile(Adam(1e-4, 0.9), 'mse') (models.py:1212-1213).
//
// This file is host code only and a CLIENT of the op-level ABI (sr_conv_plan_*, sr_wgrad_plan_*, sr_head1x1_*, ...):
// an sr_model owns the packed weights, one set of plans per (shape, pointers) and the CUDA graph that replays them.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <list>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "internal.h"

namespace {

using sr::set_cuda_error;
using sr::set_error;

constexpr int kC = 128;

#define SR_TRY(expr)               \
  do {                             \
    int rc_ = (expr);              \
    if (rc_ != SR_OK) return rc_;  \
  } while (0)
#define CU_TRY(expr, what)                                   \
  do {                                                       \
    cudaError_t e_ = (expr);                                 \
    if (e_ != cudaSuccess) return set_cuda_error(e_, what);  \
  } while (0)

struct Layer {
  char name[16];
  int k, cin, cout;
  size_t w_off, b_off;  // floats into the parameter arena
};

// 'level1', conv2d_1..conv2d_85 in Keras creation order: k3,k5,k5,k3 per 5/3 block (models.py:1253-1259), k3,k3 per
// light block (:1235-1240), the 128->3 tail last (:1199).
const std::vector<Layer>& layers() {
  static std::vector<Layer> L;
  if (!L.empty()) return L;
  std::vector<Layer> v;
  size_t off = 0;
  auto add = [&](const char* name, int k, int cin, int cout) {
    Layer l;
    snprintf(l.name, sizeof l.name, "%s", name);
    l.k = k, l.cin = cin, l.cout = cout;
    l.w_off = off;
    off += (size_t)k * k * cin * cout;
    l.b_off = off;
    off += cout;
    v.push_back(l);
  };
  add("level1", 1, 3, kC);
  int n = 0;
  char nm[16];
  auto conv = [&](int k, int cout) {
    snprintf(nm, sizeof nm, "conv2d_%d", ++n);
    add(nm, k, kC, cout);
  };
  for (int b = 0; b < 16; ++b) { conv(3, kC); conv(5, kC); conv(5, kC); conv(3, kC); }
  for (int b = 0; b < 6; ++b) { conv(3, kC); conv(3, kC); }
  for (int b = 0; b < 2; ++b) { conv(3, kC); conv(5, kC); conv(5, kC); conv(3, kC); }
  conv(3, 3);
  L.swap(v);
  return L;
}

size_t param_count() {
  const Layer& l = layers().back();
  return l.b_off + l.cout;
}

// ------------------------------------------------------------------ tiny kernels of the host-side glue
// pair_bias[j][c] = params[off[2j] + c] + params[off[2j+1] + c]: the bias of a two-source launch (5/3 block tail)
__global__ void pair_bias_kernel(const float* params, const unsigned long long* off, int npairs, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npairs * kC) {
    const int j = i / kC, c = i % kC;
    out[i] = params[off[2 * j] + c] + params[off[2 * j + 1] + c];
  }
}
// B[j = (ky*3+kx)*3+co][ci] = W[ky][kx][ci][co]: the tail's input gradient as a 1x1 conv over the im2col'ed loss
// gradient (sr_mse_tail_grad_col); rows 27..127 stay zero
__global__ void tail_colw_kernel(const float* w_hwio, float* colw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 27 * kC) {
    const int j = i / kC, ci = i % kC;
    const int tap = j / 3, co = j % 3;
    colw[(size_t)j * kC + ci] = w_hwio[((size_t)tap * kC + ci) * 3 + co];
  }
}
// dW_tail[ky][kx][ci][co] = D[ci][(ky*3+kx)*3+co], D = the k=1 filter gradient of (tail input, im2col'ed gradient)
__global__ void tail_dw_kernel(const float* d128, float* dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 9 * kC * 3) {
    const int co = i % 3, ci = (i / 3) % kC, tap = i / (3 * kC);
    dw[i] = d128[(size_t)ci * kC + tap * 3 + co];
  }
}

// ------------------------------------------------------------------ launch lists
struct Step {
  sr_conv_plan* cp = nullptr;
  sr_wgrad_plan* wp = nullptr;
  sr_conv_chain* ch = nullptr;           // many convolutions in one persistent launch (small inputs)
  std::function<int(cudaStream_t)> fn;   // anything else
  bool par_with_prev = false;            // independent of the previous step: may run beside it on the side stream
  double flops = 0;                      // tensor-core launches only
  int run(cudaStream_t st) const {
    if (cp) return sr_conv_plan_run(cp, st);
    if (wp) return sr_wgrad_plan_run(wp, st);
    if (ch) return sr_conv_chain_run(ch, st);
    return fn(st);
  }
};

struct Sequence {
  std::vector<Step> steps;
  std::vector<cudaEvent_t> events;   // fork / join pairs of the parallel steps
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
  bool ran_eager = false;
  double conv_flops = 0;
  int conv_launches = 0;
  std::vector<void*> owned_dev;      // cudaMalloc'ed by this sequence (index tables)

  ~Sequence() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    for (auto& s : steps) {
      if (s.cp) sr_conv_plan_destroy(s.cp);
      if (s.wp) sr_wgrad_plan_destroy(s.wp);
      if (s.ch) sr_conv_chain_destroy(s.ch);
    }
    for (auto e : events) cudaEventDestroy(e);
    for (auto p : owned_dev) cudaFree(p);
  }
};

struct Bump {   // carves 256-byte aligned tensors out of the caller's workspace (or just counts)
  char* base;
  size_t off = 0;
  explicit Bump(void* b) : base(reinterpret_cast<char*>(b)) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  }
};

}  // namespace

struct sr_model {
  sr_model_config cfg;
  float* params = nullptr;
  int sms = 148;
  bool tf32 = false;
  std::vector<void*> packed, packed_t;      // per layer (null for the head)
  float* pair_bias = nullptr;               // [18][128]
  unsigned long long* pair_off = nullptr;   // [18][2] bias offsets
  std::vector<std::pair<int, int>> pairs;   // layer indices of the two-source launches
  sr_pack_item* items = nullptr;            // device pack tables (forward / transposed)
  sr_pack_item* items_t = nullptr;
  unsigned long long* starts = nullptr;
  unsigned long long* starts_t = nullptr;
  int n_items = 0;
  size_t total_elems = 0, total_elems_t = 0;
  // training-only state (allocated at first use)
  void* wgrad_ws = nullptr;
  void* wgrad_ws2 = nullptr;                // second split-K scratch: two wgrad launches may run side by side
  size_t wgrad_ws_bytes = 0;
  float* tail_d128 = nullptr;
  float* tail_colw = nullptr;
  void* tail_colw_packed = nullptr;
  cudaStream_t side = nullptr, cap = nullptr;
  std::mutex mu;
  std::list<std::pair<std::string, Sequence*>> fwd_cache, train_cache;   // most recently used first
  std::vector<void*> owned;

  const float* bias_of(int li) const { return params + layers()[li].b_off; }
  const float* pair_bias_of(int a, int b) const {
    for (size_t j = 0; j < pairs.size(); ++j)
      if (pairs[j].first == a && pairs[j].second == b) return pair_bias + j * kC;
    return nullptr;
  }
  ~sr_model() {
    for (auto& e : fwd_cache) delete e.second;
    for (auto& e : train_cache) delete e.second;
    for (auto p : owned) cudaFree(p);
    if (side) cudaStreamDestroy(side);
    if (cap) cudaStreamDestroy(cap);
  }
};

namespace {

constexpr size_t kFwdCache = 24, kTrainCache = 3;

template <typename T>
int dev_alloc(sr_model* m, T** p, size_t bytes, bool zero = false) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes ? bytes : 4);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(SR_ERR_NOMEM, "sr_model: cudaMalloc failed");
  }
  if (zero) cudaMemset(q, 0, bytes);
  m->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return SR_OK;
}

int build_pack_table(sr_model* m, bool flip, std::vector<void*>* packed, sr_pack_item** items_dev,
                     unsigned long long** starts_dev, size_t* total) {
  const auto& L = layers();
  packed->assign(L.size(), nullptr);
  std::vector<sr_pack_item> items;
  std::vector<unsigned long long> starts{0};
  size_t tot = 0;
  for (size_t i = 0; i < L.size(); ++i) {
    if (L[i].cin != kC) continue;
    const size_t bytes = sr_packed_weight_bytes(L[i].k, flip ? kC : L[i].cout);
    void* dst = nullptr;
    SR_TRY(dev_alloc(m, &dst, bytes));
    (*packed)[i] = dst;
    sr_pack_item it;
    memset(&it, 0, sizeof it);
    it.hwio = m->params + L[i].w_off;
    it.dst = dst;
    it.ksize = L[i].k;
    it.cout = L[i].cout;
    it.transpose_flip = flip ? 1 : 0;
    items.push_back(it);
    tot += bytes / 2;
    starts.push_back(tot);
  }
  SR_TRY(dev_alloc(m, items_dev, items.size() * sizeof(sr_pack_item)));
  SR_TRY(dev_alloc(m, starts_dev, starts.size() * sizeof(unsigned long long)));
  CU_TRY(cudaMemcpy(*items_dev, items.data(), items.size() * sizeof(sr_pack_item), cudaMemcpyHostToDevice),
         "sr_model: pack table upload");
  CU_TRY(cudaMemcpy(*starts_dev, starts.data(), starts.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice),
         "sr_model: pack table upload");
  m->n_items = (int)items.size();
  *total = tot;
  return SR_OK;
}

int ensure_training_state(sr_model* m) {
  if (m->wgrad_ws) return SR_OK;
  if (m->tf32)
    return set_error(SR_ERR_UNSUPPORTED, "training runs on the bf16 model (dgrad / wgrad take bf16 operands); "
                                         "precision tf32 is an inference mode");
  m->wgrad_ws_bytes = sr_wgrad_workspace_bytes();
  SR_TRY(dev_alloc(m, &m->wgrad_ws, m->wgrad_ws_bytes));
  SR_TRY(dev_alloc(m, &m->wgrad_ws2, m->wgrad_ws_bytes));
  SR_TRY(dev_alloc(m, &m->tail_d128, (size_t)kC * kC * 4, true));
  SR_TRY(dev_alloc(m, &m->tail_colw, (size_t)kC * kC * 4, true));
  SR_TRY(dev_alloc(m, &m->tail_colw_packed, sr_packed_weight_bytes(1, kC)));
  SR_TRY(build_pack_table(m, true, &m->packed_t, &m->items_t, &m->starts_t, &m->total_elems_t));
  return SR_OK;
}

int refresh_transposed(sr_model* m, cudaStream_t st) {
  SR_TRY(sr_pack_conv_weights_batched(m->items_t, m->starts_t, m->n_items, m->total_elems_t, st));
  const Layer& tail = layers().back();
  tail_colw_kernel<<<(27 * kC + 255) / 256, 256, 0, st>>>(m->params + tail.w_off, m->tail_colw);
  SR_TRY(sr::check_launch("tail_colw_kernel"));
  return sr_pack_conv_weights(m->tail_colw, 1, kC, 0, m->tail_colw_packed, st);
}

int refresh(sr_model* m, cudaStream_t st) {
  const auto& L = layers();
  if (m->tf32) {
    for (size_t i = 0; i < L.size(); ++i)
      if (m->packed[i]) SR_TRY(sr_pack_conv_weights_tf32(m->params + L[i].w_off, L[i].k, L[i].cout, m->packed[i], st));
  } else {
    SR_TRY(sr_pack_conv_weights_batched(m->items, m->starts, m->n_items, m->total_elems, st));
  }
  const int np = (int)m->pairs.size();
  pair_bias_kernel<<<(np * kC + 255) / 256, 256, 0, st>>>(m->params, m->pair_off, np, m->pair_bias);
  SR_TRY(sr::check_launch("pair_bias_kernel"));
  if (m->wgrad_ws) SR_TRY(refresh_transposed(m, st));
  return SR_OK;
}

// ------------------------------------------------------------------ plan builders
struct ConvArgs {
  int nsrc = 1;
  int layer[2] = {0, 0};
  const void* in[2] = {nullptr, nullptr};
  bool flip = false;
  int NB = 0, H = 0, W = 0;
  void* out_op = nullptr;        // the operand copy: bf16, or tf32-rounded fp32 in tf32 mode
  float* out_f32 = nullptr;
  int relu = 0;
  float alpha = 1.f, beta = 0.f;
  const float* res32 = nullptr;
  const void* res16 = nullptr;
  const void* mask = nullptr;
  int cout = kC;
  bool bias = true;
  const void* wpacked_override = nullptr;   // the tail's 1x1 input-gradient weights
  int ksize_override = 0;
  const int* out_index = nullptr;
  int out_h = 0, out_w = 0;
  int comp_h = 0, comp_w = 0;
  float* colsum = nullptr;
  float colsum_scale = 0.f;
  const sr_stitch_tile* stitch_tiles = nullptr;
  uint8_t* stitch_u8 = nullptr;
  float stitch_mul = 0.f;
};

int make_desc(const sr_model* m, const ConvArgs& a, sr_conv_desc* out) {
  const auto& L = layers();
  sr_conv_desc& d = *out;
  memset(&d, 0, sizeof d);
  d.nsrc = a.nsrc;
  for (int s = 0; s < a.nsrc; ++s) {
    d.in[s] = a.in[s];
    d.wpacked[s] = a.wpacked_override ? a.wpacked_override : (a.flip ? m->packed_t : m->packed)[a.layer[s]];
    d.ksize[s] = a.ksize_override ? a.ksize_override : L[a.layer[s]].k;
  }
  d.NB = a.NB, d.H = a.H, d.W = a.W;
  d.cin = kC, d.cout = a.cout;
  d.bias = !a.bias ? nullptr : a.nsrc == 1 ? m->bias_of(a.layer[0]) : m->pair_bias_of(a.layer[0], a.layer[1]);
  if (a.bias && !d.bias) return set_error(SR_ERR_INVALID, "sr_model: unknown two-source layer pair");
  d.alpha = a.alpha, d.beta = a.beta, d.relu = a.relu;
  d.res_f32 = a.res32;
  d.out_f32 = a.out_f32;
  if (m->tf32) {
    d.precision = 1;
    d.out_tf32 = reinterpret_cast<float*>(a.out_op);
  } else {
    d.res_bf16 = (a.res16 && !a.res32) ? a.res16 : nullptr;
    d.out_bf16 = a.out_op;
  }
  d.relu_mask_bf16 = a.mask;
  d.a_mode = m->cfg.a_mode, d.nacc = m->cfg.nacc, d.pair = m->cfg.pair;
  d.out_index = a.out_index, d.out_h = a.out_h, d.out_w = a.out_w;
  d.comp_h = a.comp_h, d.comp_w = a.comp_w;
  d.colsum_f32 = a.colsum, d.colsum_scale = a.colsum_scale;
  d.stitch_tiles = a.stitch_tiles, d.stitch_u8 = a.stitch_u8, d.stitch_mul = a.stitch_mul;
  return SR_OK;
}

// When non-null, add_conv only RECORDS the convolution (descriptor + phase) instead of creating a plan: the LR stage
// of a small input becomes one chain launch (sr_conv_chain_create).
struct ChainBuilder {
  std::vector<sr_conv_desc> descs;
  std::vector<int> phase;
  int cur_phase = 0;
};

int add_conv(sr_model* m, Sequence* seq, const ConvArgs& a, sr_conv_plan_info_t* info_out = nullptr,
             ChainBuilder* chain = nullptr) {
  sr_conv_desc d;
  SR_TRY(make_desc(m, a, &d));
  if (chain) {
    chain->descs.push_back(d);
    chain->phase.push_back(chain->cur_phase);
    if (info_out) memset(info_out, 0, sizeof *info_out), info_out->grid = 1 << 20;
    return SR_OK;
  }
  Step st;
  SR_TRY(sr_conv_plan_create(&d, &st.cp));
  sr_conv_plan_info_t info;
  sr_conv_plan_info(st.cp, &info);
  st.flops = info.flops;
  seq->conv_flops += info.flops;
  seq->conv_launches += 1;
  seq->steps.push_back(std::move(st));
  if (info_out) *info_out = info;
  return SR_OK;
}

int add_fn(Sequence* seq, std::function<int(cudaStream_t)> fn) {
  Step st;
  st.fn = std::move(fn);
  seq->steps.push_back(std::move(st));
  return SR_OK;
}

struct Ext { int h, w; };

// one 5/3 block of the forward: t1 = relu(conv3_a(s)), t2 = relu(conv5_c(s)), s = 0.1*(conv5_b(t1)+conv3_d(t2)) + 0.9*s
int fwd_block53(sr_model* m, Sequence* seq, int li, void* s, float* s32, void* t1, void* t2, int NB, int H, int W,
                Ext c_out, Ext c_t1, Ext c_t2, const void* res16, ChainBuilder* chain = nullptr) {
  ConvArgs a;
  a.NB = NB, a.H = H, a.W = W;
  a.layer[0] = li, a.in[0] = s, a.out_op = t1, a.relu = 1, a.comp_h = c_t1.h, a.comp_w = c_t1.w;
  sr_conv_plan_info_t ia, ib;
  SR_TRY(add_conv(m, seq, a, &ia, chain));             // the two heads share a phase of a chain
  a.layer[0] = li + 2, a.out_op = t2, a.comp_h = c_t2.h, a.comp_w = c_t2.w;
  SR_TRY(add_conv(m, seq, a, &ib, chain));
  if (chain) chain->cur_phase += 1;
  if (!chain && m->cfg.overlap_heads && ia.grid + ib.grid <= m->sms) seq->steps.back().par_with_prev = true;
  ConvArgs f;
  f.NB = NB, f.H = H, f.W = W, f.nsrc = 2;
  f.layer[0] = li + 1, f.in[0] = t1, f.layer[1] = li + 3, f.in[1] = t2;
  f.out_op = s, f.out_f32 = s32, f.alpha = 0.1f, f.beta = 0.9f, f.res32 = s32, f.res16 = res16;
  f.comp_h = c_out.h, f.comp_w = c_out.w;
  SR_TRY(add_conv(m, seq, f, nullptr, chain));
  if (chain) chain->cur_phase += 1;
  return SR_OK;
}

int fwd_light(sr_model* m, Sequence* seq, int li, void* s, float* s32, void* t1, int NB, int H, int W, Ext c_out,
              Ext c_t1, const void* res16, ChainBuilder* chain = nullptr) {
  ConvArgs a;
  a.NB = NB, a.H = H, a.W = W;
  a.layer[0] = li, a.in[0] = s, a.out_op = t1, a.relu = 1, a.comp_h = c_t1.h, a.comp_w = c_t1.w;
  SR_TRY(add_conv(m, seq, a, nullptr, chain));
  if (chain) chain->cur_phase += 1;
  ConvArgs f;
  f.NB = NB, f.H = H, f.W = W;
  f.layer[0] = li + 1, f.in[0] = t1;
  f.out_op = s, f.out_f32 = s32, f.alpha = 0.1f, f.beta = 1.0f, f.res32 = s32, f.res16 = res16;
  f.comp_h = c_out.h, f.comp_w = c_out.w;
  SR_TRY(add_conv(m, seq, f, nullptr, chain));
  if (chain) chain->cur_phase += 1;
  return SR_OK;
}

// ------------------------------------------------------------------ running a sequence
int launch_all(sr_model* m, Sequence* seq, cudaStream_t main) {
  size_t ev = 0;
  const size_t n = seq->steps.size();
  for (size_t i = 0; i < n;) {
    if (i + 1 < n && seq->steps[i + 1].par_with_prev && m->side) {
      if (seq->events.size() < ev + 2) {
        cudaEvent_t e0, e1;
        CU_TRY(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming), "cudaEventCreate");
        CU_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming), "cudaEventCreate");
        seq->events.push_back(e0);
        seq->events.push_back(e1);
      }
      cudaEvent_t fork = seq->events[ev], join = seq->events[ev + 1];
      ev += 2;
      CU_TRY(cudaEventRecord(fork, main), "cudaEventRecord");
      CU_TRY(cudaStreamWaitEvent(m->side, fork, 0), "cudaStreamWaitEvent");
      SR_TRY(seq->steps[i + 1].run(m->side));
      SR_TRY(seq->steps[i].run(main));
      CU_TRY(cudaEventRecord(join, m->side), "cudaEventRecord");
      CU_TRY(cudaStreamWaitEvent(main, join, 0), "cudaStreamWaitEvent");
      i += 2;
    } else {
      SR_TRY(seq->steps[i].run(main));
      i += 1;
    }
  }
  return SR_OK;
}

// eager the first time; the second call captures the launches on an internal stream and from then on every call is
// one cudaGraphLaunch on the caller's stream.  A caller that is itself capturing gets the plain launches.
int run_sequence(sr_model* m, Sequence* seq, cudaStream_t st) {
  if (seq->exec) {
    CU_TRY(cudaGraphLaunch(seq->exec, st), "cudaGraphLaunch");
    return SR_OK;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
    cudaGetLastError();
    cs = cudaStreamCaptureStatusNone;
  }
  if (!m->cfg.use_graphs || !seq->ran_eager || cs != cudaStreamCaptureStatusNone) {
    SR_TRY(launch_all(m, seq, st));
    if (cs == cudaStreamCaptureStatusNone) seq->ran_eager = true;
    return SR_OK;
  }
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(m->cap, cudaStreamCaptureModeRelaxed);
  int rc = SR_OK;
  if (e == cudaSuccess) {
    rc = launch_all(m, seq, m->cap);
    e = cudaStreamEndCapture(m->cap, &graph);
  }
  if (e == cudaSuccess && rc == SR_OK && graph) {
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e == cudaSuccess) {
      seq->graph = graph;
      seq->exec = exec;
      CU_TRY(cudaGraphLaunch(seq->exec, st), "cudaGraphLaunch");
      return SR_OK;
    }
  }
  // capture is unavailable here: say so once and stay eager
  if (graph) cudaGraphDestroy(graph);
  cudaGetLastError();
  fprintf(stderr, "sr100: CUDA graph capture of a model sequence failed (%s); running eagerly\n",
          e != cudaSuccess ? cudaGetErrorString(e) : sr_last_error_string());
  m->cfg.use_graphs = 0;
  return launch_all(m, seq, st);
}

Sequence* cache_find(std::list<std::pair<std::string, Sequence*>>& cache, const std::string& key) {
  for (auto it = cache.begin(); it != cache.end(); ++it)
    if (it->first == key) {
      cache.splice(cache.begin(), cache, it);
      return cache.front().second;
    }
  return nullptr;
}

void cache_put(std::list<std::pair<std::string, Sequence*>>& cache, const std::string& key, Sequence* seq, size_t cap) {
  cache.emplace_front(key, seq);
  while (cache.size() > cap) {
    delete cache.back().second;
    cache.pop_back();
  }
}

template <typename T>
void key_add(std::string* k, const T& v) { k->append(reinterpret_cast<const char*>(&v), sizeof v); }

// ------------------------------------------------------------------ forward
struct FwdGeom {
  int n_groups;
  std::vector<int> eh, ew, n, index;
  int need_h, need_w;
  bool whole;
};

int forward_geometry(const sr_forward_desc* d, FwdGeom* g) {
  if (!d || d->NB < 1 || d->H < 1 || d->W < 1) return set_error(SR_ERR_INVALID, "sr_model_forward: empty batch");
  g->whole = d->n_groups <= 0;
  if (g->whole) {
    g->n_groups = 1;
    g->eh = {4 * d->H}, g->ew = {4 * d->W}, g->n = {d->NB};
    g->index.resize(d->NB);
    for (int i = 0; i < d->NB; ++i) g->index[i] = i;
    g->need_h = d->H, g->need_w = d->W;
    return SR_OK;
  }
  if (!d->group_eh || !d->group_ew || !d->group_n || !d->group_index)
    return set_error(SR_ERR_INVALID, "sr_model_forward: n_groups > 0 needs the four group arrays");
  g->n_groups = d->n_groups;
  int total = 0, mh = 0, mw = 0;
  for (int i = 0; i < d->n_groups; ++i) {
    const int eh = d->group_eh[i], ew = d->group_ew[i], n = d->group_n[i];
    if (n < 1 || eh < 4 || ew < 4 || eh > 4 * d->H || ew > 4 * d->W || eh % 4 || ew % 4)
      return set_error(SR_ERR_INVALID, "sr_model_forward: group extents must be multiples of 4 within the patch");
    g->eh.push_back(eh), g->ew.push_back(ew), g->n.push_back(n);
    total += n;
    mh = std::max(mh, eh), mw = std::max(mw, ew);
  }
  if (total != d->NB) return set_error(SR_ERR_INVALID, "sr_model_forward: sum(group_n) must equal NB");
  g->index.assign(d->group_index, d->group_index + total);
  for (int v : g->index)
    if (v < 0 || v >= d->NB) return set_error(SR_ERR_INVALID, "sr_model_forward: group_index out of range");
  // the bilinear reads LR cells [0, e/4] of the stream: the LR layers shrink towards that region
  g->need_h = std::min(d->H, mh / 4 + 1), g->need_w = std::min(d->W, mw / 4 + 1);
  return SR_OK;
}

struct FwdLayout {
  void *s_lr, *t1_lr, *t2_lr, *s_hr, *t1_hr, *t2_hr;
  float *s_lr32, *s_hr32;
  size_t bytes;
};

FwdLayout forward_layout(const sr_model* m, const sr_forward_desc* d, const FwdGeom& g, void* base) {
  const size_t ops = m->tf32 ? 4 : 2;
  const size_t lr = (size_t)d->NB * d->H * d->W * kC;
  size_t hr = 0;
  for (int i = 0; i < g.n_groups; ++i) hr = std::max(hr, (size_t)g.n[i] * g.eh[i] * g.ew[i] * kC);
  Bump b(base);
  FwdLayout l;
  l.s_lr = b.take(lr * ops), l.t1_lr = b.take(lr * ops), l.t2_lr = b.take(lr * ops);
  l.s_lr32 = m->cfg.stream_lr_fp32 ? reinterpret_cast<float*>(b.take(lr * 4)) : nullptr;
  l.s_hr = b.take(hr * ops), l.t1_hr = b.take(hr * ops), l.t2_hr = b.take(hr * ops);
  l.s_hr32 = m->cfg.stream_hr_fp32 ? reinterpret_cast<float*>(b.take(hr * 4)) : nullptr;
  l.bytes = b.off;
  return l;
}

// Compute extents of every conv of the 22 LR blocks, given the region `need` of the final LR stream that is read
// afterwards (walking backwards: a light block grows the needed region by 2 pixels per axis, a 5/3 block by 3).
void lr_extents(Ext need, Ext full, Ext (*out)[3]) {
  auto clip = [&](Ext e, int add) { return Ext{std::min(full.h, e.h + add), std::min(full.w, e.w + add)}; };
  Ext e{std::min(full.h, need.h), std::min(full.w, need.w)};
  for (int b = 21; b >= 0; --b) {
    if (b >= 16) {
      out[b][0] = e, out[b][1] = clip(e, 1), out[b][2] = Ext{0, 0};
      e = clip(e, 2);
    } else {
      out[b][0] = e, out[b][1] = clip(e, 2), out[b][2] = clip(e, 1);
      e = clip(e, 3);
    }
  }
}

int build_forward(sr_model* m, const sr_forward_desc* d, const FwdGeom& g, Sequence* seq) {
  const FwdLayout l = forward_layout(m, d, g, d->workspace);
  const int NB = d->NB, H = d->H, W = d->W;
  const size_t npix = (size_t)NB * H * W;
  const float* head_w = m->params + layers()[0].w_off;
  const float* head_b = m->params + layers()[0].b_off;
  const float* x = d->x;
  if (m->tf32) {
    float* s32 = l.s_lr32;
    float* s = reinterpret_cast<float*>(l.s_lr);
    add_fn(seq, [=](cudaStream_t st) { return sr_head1x1_fwd(x, head_w, head_b, npix, nullptr, s32, st); });
    add_fn(seq, [=](cudaStream_t st) { return sr_round_tf32(s32, npix * kC, s, st); });
  } else {
    void* s = l.s_lr;
    float* s32 = l.s_lr32;
    add_fn(seq, [=](cudaStream_t st) { return sr_head1x1_fwd(x, head_w, head_b, npix, s, s32, st); });
  }
  Ext exts[22][3];
  lr_extents(Ext{g.need_h, g.need_w}, Ext{H, W}, exts);
  // Opt-in (sr_model_config.chain_lr): a small LR stage (every CTA pair gets about one 128-position tile per layer:
  // a single 128 x 128 patch, BASELINE config 1) as ONE persistent chain launch instead of 60 launches.  Bit-identical,
  // but measured SLOWER than the per-layer launches under a CUDA graph (the grid barrier + the exposed
  // strip burst and epilogue of a one-tile phase cost more than the launch boundary they replace), so it is off by
  // default.
  const bool small_lr = npix <= (size_t)m->sms * 128;
  bool chained = false;
  int li = 1;
  if (m->cfg.chain_lr && small_lr && !m->tf32 && l.s_lr32 && m->cfg.a_mode == 0 && m->cfg.pair == 1 && m->cfg.nacc == 2) {
    ChainBuilder cb;
    int lj = 1;
    int rc = SR_OK;
    for (int b = 0; b < 16 && rc == SR_OK; ++b, lj += 4)
      rc = fwd_block53(m, seq, lj, l.s_lr, l.s_lr32, l.t1_lr, l.t2_lr, NB, H, W, exts[b][0], exts[b][1], exts[b][2],
                       l.s_lr, &cb);
    for (int b = 16; b < 22 && rc == SR_OK; ++b, lj += 2)
      rc = fwd_light(m, seq, lj, l.s_lr, l.s_lr32, l.t1_lr, NB, H, W, exts[b][0], exts[b][1], l.s_lr, &cb);
    SR_TRY(rc);
    Step st;
    rc = sr_conv_chain_create(cb.descs.data(), cb.phase.data(), (int)cb.descs.size(), &st.ch);
    if (rc == SR_OK) {
      sr_conv_plan_info_t info;
      sr_conv_chain_info(st.ch, &info);
      st.flops = info.flops;
      seq->conv_flops += info.flops;
      seq->conv_launches += 1;
      seq->steps.push_back(std::move(st));
      chained = true;
      li = lj;
    } else if (rc != SR_ERR_UNSUPPORTED) {
      return rc;
    }
  }
  if (!chained) {
    for (int b = 0; b < 16; ++b, li += 4)
      SR_TRY(fwd_block53(m, seq, li, l.s_lr, l.s_lr32, l.t1_lr, l.t2_lr, NB, H, W, exts[b][0], exts[b][1], exts[b][2],
                         l.s_lr));
    for (int b = 16; b < 22; ++b, li += 2)
      SR_TRY(fwd_light(m, seq, li, l.s_lr, l.s_lr32, l.t1_lr, NB, H, W, exts[b][0], exts[b][1], l.s_lr));
  }
  const int first_hr = li;
  // the index tables of all groups: one device array, uploaded once
  int* idx_dev = nullptr;
  CU_TRY(cudaMalloc(reinterpret_cast<void**>(&idx_dev), g.index.size() * sizeof(int)), "sr_model: cudaMalloc(index)");
  seq->owned_dev.push_back(idx_dev);
  CU_TRY(cudaMemcpy(idx_dev, g.index.data(), g.index.size() * sizeof(int), cudaMemcpyHostToDevice),
         "sr_model: index upload");
  const void* src = l.s_lr32 ? static_cast<const void*>(l.s_lr32) : l.s_lr;
  const int src_is_bf16 = l.s_lr32 ? 0 : 1;
  int off = 0;
  for (int gi = 0; gi < g.n_groups; ++gi) {
    const int n = g.n[gi], eh = g.eh[gi], ew = g.ew[gi];
    const int* idx = idx_dev + off;
    off += n;
    if (m->tf32) {
      float* s32 = l.s_hr32;
      float* s = reinterpret_cast<float*>(l.s_hr);
      add_fn(seq, [=](cudaStream_t st) {
        return sr_bilinear4_crop_fwd(src, 0, idx, n, H, W, kC, eh, ew, nullptr, s32, st);
      });
      add_fn(seq, [=](cudaStream_t st) { return sr_round_tf32(s32, (size_t)n * eh * ew * kC, s, st); });
    } else {
      void* s = l.s_hr;
      float* s32 = l.s_hr32;
      add_fn(seq, [=](cudaStream_t st) {
        return sr_bilinear4_crop_fwd(src, src_is_bf16, idx, n, H, W, kC, eh, ew, s, s32, st);
      });
    }
    // a cropped extent e (< 4H) was chosen as >= (last surviving pixel + 1) + 7: the tail only has to be right on
    // [0, e-7), the second block on e-6 (its t1 / t2 on e-4 / e-5), the first block on e-3 (t1 / t2 on e-1 / e-2)
    auto cut = [&](int k) { return Ext{eh < 4 * H ? eh - k : eh, ew < 4 * W ? ew - k : ew}; };
    li = first_hr;
    SR_TRY(fwd_block53(m, seq, li, l.s_hr, l.s_hr32, l.t1_hr, l.t2_hr, n, eh, ew, cut(3), cut(1), cut(2), l.s_hr));
    li += 4;
    SR_TRY(fwd_block53(m, seq, li, l.s_hr, l.s_hr32, l.t1_hr, l.t2_hr, n, eh, ew, cut(6), cut(4), cut(5), l.s_hr));
    li += 4;
    ConvArgs t;
    t.NB = n, t.H = eh, t.W = ew;
    t.layer[0] = li, t.in[0] = l.s_hr, t.out_f32 = d->out, t.relu = 1, t.cout = 3;
    t.out_index = idx, t.out_h = 4 * H, t.out_w = 4 * W;
    t.stitch_tiles = d->stitch_tiles, t.stitch_u8 = d->stitch_u8, t.stitch_mul = d->stitch_mul;
    const Ext c7 = cut(7);
    t.comp_h = c7.h, t.comp_w = c7.w;
    SR_TRY(add_conv(m, seq, t));
  }
  return SR_OK;
}

int get_forward(sr_model* m, const sr_forward_desc* d, Sequence** out) {
  FwdGeom g;
  SR_TRY(forward_geometry(d, &g));
  if (!d->x || !d->workspace || (!d->out && !d->stitch_u8))
    return set_error(SR_ERR_INVALID, "sr_model_forward: null tensor / workspace");
  if (d->stitch_u8 && !d->stitch_tiles) return set_error(SR_ERR_INVALID, "sr_model_forward: stitch_u8 needs stitch_tiles");
  const FwdLayout l = forward_layout(m, d, g, nullptr);
  if (d->workspace_bytes < l.bytes)
    return set_error(SR_ERR_INVALID, "sr_model_forward: workspace too small (see sr_model_forward_workspace_bytes)");
  std::string key;
  key_add(&key, d->NB), key_add(&key, d->H), key_add(&key, d->W);
  key_add(&key, d->x), key_add(&key, d->out), key_add(&key, d->workspace);
  key_add(&key, d->stitch_tiles), key_add(&key, d->stitch_u8), key_add(&key, d->stitch_mul);
  key_add(&key, g.n_groups);
  for (int v : g.eh) key_add(&key, v);
  for (int v : g.ew) key_add(&key, v);
  for (int v : g.n) key_add(&key, v);
  for (int v : g.index) key_add(&key, v);
  Sequence* seq = cache_find(m->fwd_cache, key);
  if (!seq) {
    seq = new (std::nothrow) Sequence();
    if (!seq) return set_error(SR_ERR_NOMEM, "out of host memory");
    const int rc = build_forward(m, d, g, seq);
    if (rc != SR_OK) {
      delete seq;
      return rc;
    }
    cache_put(m->fwd_cache, key, seq, kFwdCache);
  }
  *out = seq;
  return SR_OK;
}

// ------------------------------------------------------------------ training
struct TrainLayout {
  void *S[23], *T1[22], *T2[16], *SH[3], *TH1[2], *TH2[2];
  float *s32, *gs32, *gsh32, *pred;
  void *gs, *gt1, *gt2, *gsh, *gth1, *gth2;
  size_t bytes;
};

TrainLayout train_layout(int NB, int H, int W, void* base, bool own_pred) {
  const size_t lr = (size_t)NB * H * W * kC, hr = lr * 16;
  Bump b(base);
  TrainLayout l;
  for (auto& p : l.S) p = b.take(lr * 2);
  for (auto& p : l.T1) p = b.take(lr * 2);
  for (auto& p : l.T2) p = b.take(lr * 2);
  l.s32 = reinterpret_cast<float*>(b.take(lr * 4));
  for (auto& p : l.SH) p = b.take(hr * 2);
  for (auto& p : l.TH1) p = b.take(hr * 2);
  for (auto& p : l.TH2) p = b.take(hr * 2);
  l.gs = b.take(lr * 2), l.gs32 = reinterpret_cast<float*>(b.take(lr * 4));
  l.gt1 = b.take(lr * 2), l.gt2 = b.take(lr * 2);
  l.gsh = b.take(hr * 2), l.gsh32 = reinterpret_cast<float*>(b.take(hr * 4));
  l.gth1 = b.take(hr * 2), l.gth2 = b.take(hr * 2);
  l.pred = own_pred ? reinterpret_cast<float*>(b.take((size_t)NB * 16 * H * W * 3 * 4)) : nullptr;
  l.bytes = b.off;
  return l;
}

struct Block {
  bool is53;
  int li;
  void *x, *t1, *t2, *y;
  int NB, H, W;
  bool f32s;
};

int add_wgrad(sr_model* m, Sequence* seq, const void* x, const void* g, int NB, int H, int W, int ksize, float scale,
              float* dw, bool second_ws = false) {
  sr_wgrad_desc d;
  memset(&d, 0, sizeof d);
  d.x_bf16 = x, d.g_bf16 = g;
  d.NB = NB, d.H = H, d.W = W;
  d.ksize = ksize, d.scale = scale, d.accumulate = 0;
  d.dw_hwio = dw, d.workspace = second_ws ? m->wgrad_ws2 : m->wgrad_ws, d.workspace_bytes = m->wgrad_ws_bytes;
  Step st;
  SR_TRY(sr_wgrad_plan_create(&d, &st.wp));
  sr_wgrad_plan_info_t info;
  sr_wgrad_plan_info(st.wp, &info);
  st.flops = info.flops;
  seq->conv_flops += info.flops;
  seq->conv_launches += 1;
  seq->steps.push_back(std::move(st));
  return SR_OK;
}

int build_train(sr_model* m, const sr_train_desc* d, Sequence* seq) {
  const auto& L = layers();
  const int NB = d->NB, H = d->H, W = d->W, HH = 4 * H, WW = 4 * W;
  const TrainLayout l = train_layout(NB, H, W, d->workspace, d->pred == nullptr);
  float* pred = d->pred ? d->pred : l.pred;
  float* grads = d->grads;
  double* loss_sum = d->loss_sum;
  const size_t npix = (size_t)NB * H * W, npix_hr = npix * 16;
  auto gw = [&](int li) { return grads + L[li].w_off; };
  auto gb = [&](int li) { return grads + L[li].b_off; };
  const size_t nparams = param_count();

  add_fn(seq, [=](cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(grads, 0, nparams * 4, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(loss_sum, 0, sizeof(double), st);
    return e == cudaSuccess ? SR_OK : set_cuda_error(e, "sr_model: zero gradients");
  });
  // ---------------------------------------------------------------- forward, keeping what the backward reads
  {
    const float* x = d->x;
    const float* hw = m->params + L[0].w_off;
    const float* hb = m->params + L[0].b_off;
    void* s0 = l.S[0];
    float* s32 = l.s32;
    add_fn(seq, [=](cudaStream_t st) { return sr_head1x1_fwd(x, hw, hb, npix, s0, s32, st); });
  }
  std::vector<Block> blocks;
  int li = 1;
  for (int b = 0; b < 16; ++b, li += 4)
    blocks.push_back(Block{true, li, l.S[b], l.T1[b], l.T2[b], l.S[b + 1], NB, H, W, true});
  for (int b = 0; b < 6; ++b, li += 2)
    blocks.push_back(Block{false, li, l.S[16 + b], l.T1[16 + b], nullptr, l.S[17 + b], NB, H, W, true});
  for (int b = 0; b < 2; ++b, li += 4)
    blocks.push_back(Block{true, li, l.SH[b], l.TH1[b], l.TH2[b], l.SH[b + 1], NB, HH, WW, false});
  const int tail = li;
  const bool par = m->cfg.overlap_train != 0;   // independent launches pair up on two streams
  const Ext none{0, 0};
  auto fwd_block = [&](const Block& k) -> int {
    float* r32 = k.f32s ? l.s32 : nullptr;
    ConvArgs a;
    a.NB = k.NB, a.H = k.H, a.W = k.W;
    a.layer[0] = k.li, a.in[0] = k.x, a.out_op = k.t1, a.relu = 1;
    SR_TRY(add_conv(m, seq, a));
    ConvArgs f;
    f.NB = k.NB, f.H = k.H, f.W = k.W;
    f.out_op = k.y, f.out_f32 = r32, f.alpha = 0.1f, f.res32 = r32, f.res16 = k.x;
    if (k.is53) {
      a.layer[0] = k.li + 2, a.out_op = k.t2;
      SR_TRY(add_conv(m, seq, a));
      if (par) seq->steps.back().par_with_prev = true;
      f.nsrc = 2, f.layer[0] = k.li + 1, f.in[0] = k.t1, f.layer[1] = k.li + 3, f.in[1] = k.t2, f.beta = 0.9f;
    } else {
      f.layer[0] = k.li + 1, f.in[0] = k.t1, f.beta = 1.0f;
    }
    return add_conv(m, seq, f);
  };
  (void)none;
  for (int b = 0; b < 22; ++b) SR_TRY(fwd_block(blocks[b]));
  {
    const float* s32 = l.s32;
    void* sh0 = l.SH[0];
    add_fn(seq, [=](cudaStream_t st) { return sr_bilinear4_fwd(s32, 0, NB, H, W, kC, sh0, nullptr, st); });
  }
  for (int b = 22; b < 24; ++b) SR_TRY(fwd_block(blocks[b]));
  {
    ConvArgs t;
    t.NB = NB, t.H = HH, t.W = WW, t.layer[0] = tail, t.in[0] = l.SH[2], t.out_f32 = pred, t.relu = 1, t.cout = 3;
    SR_TRY(add_conv(m, seq, t));
  }
  // ---------------------------------------------------------------- backward
  const size_t n_local = npix_hr * 3;
  {
    // tail conv (128 -> 3, 3x3): the loss gradient is written as the im2col of the tail's backward (27 channels of a
    // 128-channel bf16 tensor), which turns both tail gradients into 1x1 problems for the tensor-core kernels
    const float* y = d->y;
    void* gth1 = l.gth1;
    float* tb = gb(tail);
    add_fn(seq, [=](cudaStream_t st) {
      return sr_mse_tail_grad_col(pred, y, NB, HH, WW, n_local, gth1, loss_sum, tb, st);
    });
    SR_TRY(add_wgrad(m, seq, l.SH[2], l.gth1, NB, HH, WW, 1, 1.0f, m->tail_d128));   // D[ci][j], j = (ky,kx,co)
    seq->steps.back().flops = 2.0 * npix_hr * 27 * kC;
    const float* d128 = m->tail_d128;
    float* tw = gw(tail);
    add_fn(seq, [=](cudaStream_t st) {
      tail_dw_kernel<<<(9 * kC * 3 + 255) / 256, 256, 0, st>>>(d128, tw);
      return sr::check_launch("tail_dw_kernel");
    });
    ConvArgs a;   // dgrad: 1x1 conv with B[j][ci] = W[ky][kx][ci][co]
    a.NB = NB, a.H = HH, a.W = WW, a.in[0] = l.gth1, a.wpacked_override = m->tail_colw_packed, a.ksize_override = 1;
    a.out_op = l.gsh, a.bias = false;
    SR_TRY(add_conv(m, seq, a));
    seq->steps.back().flops = 2.0 * npix_hr * 27 * kC;
  }
  const bool fuse_cs = m->cfg.nacc == 2 && m->cfg.a_mode == 0 && m->cfg.fused_colsum != 0;
  struct GB { float* db; float scale; float* twin; };
  auto g_bias = [&](const Block& k) {   // the 0.1-scaled block tail(s) whose output gradient is the block's incoming g
    return k.is53 ? GB{gb(k.li + 1), 0.1f, gb(k.li + 3)} : GB{gb(k.li + 1), 0.1f, nullptr};
  };
  auto colsum = [&](const void* g, const Block& k, float* db, float scale) {
    const size_t np = (size_t)k.NB * k.H * k.W;
    add_fn(seq, [=](cudaStream_t st) { return sr_colsum_bf16(g, np, scale, db, st); });
  };
  auto bwd_block = [&](const Block& k, bool g_summed, const Block* nxt, bool last_hr, bool* next_summed) -> int {
    void* g = k.f32s ? l.gs : l.gsh;
    float* g32 = k.f32s ? l.gs32 : nullptr;
    void* a1 = k.f32s ? l.gt1 : l.gth1;
    void* a2 = k.f32s ? l.gt2 : l.gth2;
    float* o32 = k.f32s ? g32 : (last_hr ? l.gsh32 : nullptr);
    const GB gbk = g_bias(k);
    // the launch that writes the next block's g can carry its column sums only if g stays bf16-resident in the same
    // buffer (not across the HR -> LR boundary, where g goes through the bilinear adjoint)
    const bool cs_next = fuse_cs && nxt && nxt->f32s == k.f32s;
    const GB gbn = cs_next ? g_bias(*nxt) : GB{nullptr, 0.f, nullptr};
    auto dgrad_masked = [&](int layer, void* out, const void* mask, int colsum_layer) {
      ConvArgs a;
      a.NB = k.NB, a.H = k.H, a.W = k.W, a.flip = true, a.bias = false;
      a.layer[0] = layer, a.in[0] = g, a.out_op = out, a.alpha = 0.1f, a.mask = mask;
      if (fuse_cs) a.colsum = gb(colsum_layer), a.colsum_scale = 1.0f;
      return add_conv(m, seq, a);
    };
    const int na = k.li, nb = k.li + 1, nc = k.li + 2, nd = k.li + 3;
    if (k.is53) {
      SR_TRY(dgrad_masked(nb, a1, k.t1, na));
      SR_TRY(dgrad_masked(nd, a2, k.t2, nc));
      if (par) seq->steps.back().par_with_prev = true;
      SR_TRY(add_wgrad(m, seq, k.t1, g, k.NB, k.H, k.W, L[nb].k, 0.1f, gw(nb)));
      SR_TRY(add_wgrad(m, seq, k.t2, g, k.NB, k.H, k.W, L[nd].k, 0.1f, gw(nd), par));
      if (par) seq->steps.back().par_with_prev = true;
      if (!g_summed) colsum(g, k, gbk.db, 0.1f);
      {
        float* twin = gbk.twin;
        const float* db = gbk.db;
        add_fn(seq, [=](cudaStream_t st) {
          cudaError_t e = cudaMemcpyAsync(twin, db, kC * 4, cudaMemcpyDeviceToDevice, st);
          return e == cudaSuccess ? SR_OK : set_cuda_error(e, "sr_model: bias gradient copy");
        });
      }
      ConvArgs f;
      f.NB = k.NB, f.H = k.H, f.W = k.W, f.flip = true, f.bias = false, f.nsrc = 2;
      f.layer[0] = na, f.in[0] = a1, f.layer[1] = nc, f.in[1] = a2;
      f.out_op = g, f.out_f32 = o32, f.alpha = 1.0f, f.beta = 0.9f, f.res32 = g32, f.res16 = g;
      if (cs_next) f.colsum = gbn.db, f.colsum_scale = gbn.scale;
      SR_TRY(add_conv(m, seq, f));
      SR_TRY(add_wgrad(m, seq, k.x, a1, k.NB, k.H, k.W, L[na].k, 1.0f, gw(na)));
      SR_TRY(add_wgrad(m, seq, k.x, a2, k.NB, k.H, k.W, L[nc].k, 1.0f, gw(nc), par));
      if (par) seq->steps.back().par_with_prev = true;
      if (!fuse_cs) {
        colsum(a1, k, gb(na), 1.0f);
        colsum(a2, k, gb(nc), 1.0f);
      }
    } else {
      SR_TRY(dgrad_masked(nb, a1, k.t1, na));
      SR_TRY(add_wgrad(m, seq, k.t1, g, k.NB, k.H, k.W, L[nb].k, 0.1f, gw(nb)));
      if (!g_summed) colsum(g, k, gbk.db, 0.1f);
      ConvArgs f;
      f.NB = k.NB, f.H = k.H, f.W = k.W, f.flip = true, f.bias = false;
      f.layer[0] = na, f.in[0] = a1;
      f.out_op = g, f.out_f32 = o32, f.alpha = 1.0f, f.beta = 1.0f, f.res32 = g32, f.res16 = g;
      if (cs_next) f.colsum = gbn.db, f.colsum_scale = gbn.scale;
      SR_TRY(add_conv(m, seq, f));
      SR_TRY(add_wgrad(m, seq, k.x, a1, k.NB, k.H, k.W, L[na].k, 1.0f, gw(na)));
      if (!fuse_cs) colsum(a1, k, gb(na), 1.0f);
    }
    *next_summed = cs_next;
    return SR_OK;
  };
  bool done = false;
  SR_TRY(bwd_block(blocks[23], false, &blocks[22], false, &done));
  bool dummy = false;
  SR_TRY(bwd_block(blocks[22], done, nullptr, true, &dummy));
  {
    const float* gsh32 = l.gsh32;
    float* gs32 = l.gs32;
    void* gs = l.gs;
    add_fn(seq, [=](cudaStream_t st) { return sr_bilinear4_bwd(gsh32, NB, H, W, kC, gs32, st); });
    add_fn(seq, [=](cudaStream_t st) { return sr_cast_f32_to_bf16(gs32, npix * kC, gs, st); });
  }
  done = false;
  for (int bi = 21; bi >= 0; --bi) SR_TRY(bwd_block(blocks[bi], done, bi > 0 ? &blocks[bi - 1] : nullptr, false, &done));
  {
    const float* x = d->x;
    const void* s0 = l.S[0];
    const float* gs32 = l.gs32;
    float* hw = gw(0);
    float* hb = gb(0);
    add_fn(seq, [=](cudaStream_t st) { return sr_head1x1_bwd(x, s0, gs32, nullptr, npix, hw, hb, st); });
  }
  return SR_OK;
}

int get_train(sr_model* m, const sr_train_desc* d, Sequence** out) {
  if (!d || d->NB < 1 || d->H < 1 || d->W < 1) return set_error(SR_ERR_INVALID, "sr_model_forward_backward: empty batch");
  if (!d->x || !d->y || !d->grads || !d->loss_sum || !d->workspace)
    return set_error(SR_ERR_INVALID, "sr_model_forward_backward: null tensor / workspace");
  SR_TRY(ensure_training_state(m));
  const TrainLayout l = train_layout(d->NB, d->H, d->W, nullptr, d->pred == nullptr);
  if (d->workspace_bytes < l.bytes)
    return set_error(SR_ERR_INVALID, "sr_model_forward_backward: workspace too small (see sr_model_train_workspace_bytes)");
  std::string key;
  key_add(&key, d->NB), key_add(&key, d->H), key_add(&key, d->W);
  key_add(&key, d->x), key_add(&key, d->y), key_add(&key, d->grads), key_add(&key, d->loss_sum);
  key_add(&key, d->pred), key_add(&key, d->workspace);
  Sequence* seq = cache_find(m->train_cache, key);
  if (!seq) {
    seq = new (std::nothrow) Sequence();
    if (!seq) return set_error(SR_ERR_NOMEM, "out of host memory");
    // the transposed weights must be current before the first backward that uses them
    int rc = build_train(m, d, seq);
    if (rc != SR_OK) {
      delete seq;
      return rc;
    }
    cache_put(m->train_cache, key, seq, kTrainCache);
  }
  *out = seq;
  return SR_OK;
}

void fill_info(const sr_model* m, const Sequence* seq, sr_model_run_info* info) {
  info->conv_flops = 0;
  for (const auto& s : seq->steps) info->conv_flops += s.flops;
  info->launches = (int)seq->steps.size();
  info->conv_launches = seq->conv_launches;
  info->graph_replay = (seq->exec || (m->cfg.use_graphs && seq->ran_eager)) ? 1 : 0;
}

}  // namespace

// ====================================================================== C ABI
extern "C" void sr_model_default_config(sr_model_config* c) {
  if (!c) return;
  memset(c, 0, sizeof *c);
  c->precision = 0, c->stream_lr_fp32 = 1, c->stream_hr_fp32 = 0;
  c->a_mode = 0, c->nacc = 2, c->pair = 1, c->use_graphs = 1, c->overlap_heads = 1, c->fused_colsum = 1;
  c->chain_lr = 0;
}

extern "C" int sr_model_num_layers(void) { return (int)layers().size(); }
extern "C" size_t sr_model_param_count(void) { return param_count(); }

extern "C" int sr_model_layer(int index, char* name, int* ksize, int* cin, int* cout, size_t* kernel_offset,
                              size_t* bias_offset) {
  const auto& L = layers();
  if (index < 0 || index >= (int)L.size()) return set_error(SR_ERR_INVALID, "sr_model_layer: index out of range");
  const Layer& l = L[index];
  if (name) snprintf(name, 16, "%s", l.name);
  if (ksize) *ksize = l.k;
  if (cin) *cin = l.cin;
  if (cout) *cout = l.cout;
  if (kernel_offset) *kernel_offset = l.w_off;
  if (bias_offset) *bias_offset = l.b_off;
  return SR_OK;
}

extern "C" int sr_model_create(float* params, const sr_model_config* cfg, sr_model** out) {
  if (!params || !out) return set_error(SR_ERR_INVALID, "sr_model_create: null argument");
  if (!sr_device_supported()) return set_error(SR_ERR_UNSUPPORTED, "sr_model_create: the current device is not sm_100");
  sr_model* m = new (std::nothrow) sr_model();
  if (!m) return set_error(SR_ERR_NOMEM, "out of host memory");
  if (cfg) m->cfg = *cfg; else sr_model_default_config(&m->cfg);
  m->params = params;
  m->tf32 = m->cfg.precision == 1;
  if (m->cfg.precision != 0 && m->cfg.precision != 1) {
    delete m;
    return set_error(SR_ERR_INVALID, "sr_model_create: precision must be 0 (bf16) or 1 (tf32)");
  }
  if (m->tf32) m->cfg.stream_lr_fp32 = m->cfg.stream_hr_fp32 = 1, m->cfg.a_mode = 0, m->cfg.nacc = 2;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&m->sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) m->sms = 148;
  auto fail = [&](int rc) {
    delete m;
    return rc;
  };
  if (cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&m->cap, cudaStreamNonBlocking) != cudaSuccess)
    return fail(set_error(SR_ERR_CUDA, "sr_model_create: cudaStreamCreate failed"));
  const auto& L = layers();
  int rc;
  if (m->tf32) {
    m->packed.assign(L.size(), nullptr);
    for (size_t i = 0; i < L.size(); ++i)
      if (L[i].cin == kC)
        if ((rc = dev_alloc(m, &m->packed[i], sr_packed_weight_bytes_tf32(L[i].k, L[i].cout))) != SR_OK) return fail(rc);
  } else if ((rc = build_pack_table(m, false, &m->packed, &m->items, &m->starts, &m->total_elems)) != SR_OK) {
    return fail(rc);
  }
  // the 18 two-source launches (fused tails of the 5/3 blocks) add two biases
  std::vector<unsigned long long> off;
  auto pair_at = [&](int li) {
    m->pairs.emplace_back(li + 1, li + 3);
    off.push_back(L[li + 1].b_off);
    off.push_back(L[li + 3].b_off);
  };
  for (int b = 0; b < 16; ++b) pair_at(1 + 4 * b);
  for (int b = 0; b < 2; ++b) pair_at(1 + 64 + 12 + 4 * b);
  if ((rc = dev_alloc(m, &m->pair_bias, m->pairs.size() * kC * 4, true)) != SR_OK) return fail(rc);
  if ((rc = dev_alloc(m, &m->pair_off, off.size() * sizeof(unsigned long long))) != SR_OK) return fail(rc);
  if (cudaMemcpy(m->pair_off, off.data(), off.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(set_error(SR_ERR_CUDA, "sr_model_create: upload failed"));
  if ((rc = refresh(m, nullptr)) != SR_OK) return fail(rc);
  if (cudaStreamSynchronize(nullptr) != cudaSuccess) return fail(set_error(SR_ERR_CUDA, "sr_model_create: pack failed"));
  *out = m;
  return SR_OK;
}

extern "C" void sr_model_destroy(sr_model* m) { delete m; }

extern "C" int sr_model_refresh(sr_model* m, void* stream) {
  if (!m) return set_error(SR_ERR_INVALID, "sr_model_refresh: null model");
  std::lock_guard<std::mutex> lk(m->mu);
  return refresh(m, sr::as_stream(stream));
}

extern "C" size_t sr_model_forward_workspace_bytes(const sr_model* m, const sr_forward_desc* d) {
  FwdGeom g;
  if (!m || forward_geometry(d, &g) != SR_OK) return 0;
  return forward_layout(m, d, g, nullptr).bytes;
}

extern "C" int sr_model_forward(sr_model* m, const sr_forward_desc* d, void* stream) {
  if (!m) return set_error(SR_ERR_INVALID, "sr_model_forward: null model");
  std::lock_guard<std::mutex> lk(m->mu);
  Sequence* seq = nullptr;
  SR_TRY(get_forward(m, d, &seq));
  return run_sequence(m, seq, sr::as_stream(stream));
}

extern "C" int sr_model_forward_info(sr_model* m, const sr_forward_desc* d, sr_model_run_info* info) {
  if (!m || !info) return set_error(SR_ERR_INVALID, "sr_model_forward_info: null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  Sequence* seq = nullptr;
  SR_TRY(get_forward(m, d, &seq));
  fill_info(m, seq, info);
  return SR_OK;
}

extern "C" int sr_model_forward_timed(sr_model* m, const sr_forward_desc* d, void* stream, float* ms, double* flops,
                                      int max_records, int* n) {
  if (!m || !n) return set_error(SR_ERR_INVALID, "sr_model_forward_timed: null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  Sequence* seq = nullptr;
  SR_TRY(get_forward(m, d, &seq));
  cudaStream_t st = sr::as_stream(stream);
  const int count = (int)seq->steps.size();
  std::vector<cudaEvent_t> ev(2 * (size_t)count);
  for (auto& e : ev) CU_TRY(cudaEventCreate(&e), "cudaEventCreate");
  int rc = SR_OK;
  for (int i = 0; i < count && rc == SR_OK; ++i) {
    cudaEventRecord(ev[2 * i], st);
    rc = seq->steps[i].run(st);
    cudaEventRecord(ev[2 * i + 1], st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  for (int i = 0; i < count && i < max_records && rc == SR_OK && e == cudaSuccess; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
    if (ms) ms[i] = t;
    if (flops) flops[i] = seq->steps[i].flops;
  }
  for (auto& x : ev) cudaEventDestroy(x);
  *n = count;
  if (rc != SR_OK) return rc;
  if (e != cudaSuccess) return set_cuda_error(e, "sr_model_forward_timed");
  return SR_OK;
}

extern "C" size_t sr_model_train_workspace_bytes(const sr_model* m, int NB, int H, int W) {
  if (!m || NB < 1 || H < 1 || W < 1) return 0;
  return train_layout(NB, H, W, nullptr, true).bytes;
}

extern "C" int sr_model_forward_backward(sr_model* m, const sr_train_desc* d, void* stream) {
  if (!m) return set_error(SR_ERR_INVALID, "sr_model_forward_backward: null model");
  std::lock_guard<std::mutex> lk(m->mu);
  const bool first = m->wgrad_ws == nullptr;
  Sequence* seq = nullptr;
  SR_TRY(get_train(m, d, &seq));
  if (first) SR_TRY(refresh_transposed(m, sr::as_stream(stream)));
  return run_sequence(m, seq, sr::as_stream(stream));
}

extern "C" int sr_model_train_info(sr_model* m, const sr_train_desc* d, sr_model_run_info* info) {
  if (!m || !info) return set_error(SR_ERR_INVALID, "sr_model_train_info: null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  const bool first = m->wgrad_ws == nullptr;
  Sequence* seq = nullptr;
  SR_TRY(get_train(m, d, &seq));
  if (first) SR_TRY(refresh_transposed(m, nullptr));
  fill_info(m, seq, info);
  return SR_OK;
}

extern "C" int sr_model_apply_gradients(sr_model* m, const float* grads, float* mom, float* vel, int t, float lr,
                                        float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!m || !grads || !mom || !vel) return set_error(SR_ERR_INVALID, "sr_model_apply_gradients: null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  SR_TRY(sr_adam_step(m->params, grads, mom, vel, param_count(), lr, beta1, beta2, eps, t, grad_scale, stream));
  return refresh(m, sr::as_stream(stream));
}

extern "C" int sr_model_apply_gradients_exchange(sr_model* m, sr_exchange* ex, float* mom, float* vel, int t, float lr,
                                                 float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!m || !ex || !mom || !vel) return set_error(SR_ERR_INVALID, "sr_model_apply_gradients_exchange: null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  SR_TRY(sr_exchange_adam_step(ex, mom, vel, t, lr, beta1, beta2, eps, grad_scale, 0, stream));
  return refresh(m, sr::as_stream(stream));
}

extern "C" int sr_model_train_step(sr_model* m, const sr_train_desc* d, float* mom, float* vel, int t, float lr,
                                   float beta1, float beta2, float eps, void* stream) {
  SR_TRY(sr_model_forward_backward(m, d, stream));
  return sr_model_apply_gradients(m, d->grads, mom, vel, t, lr, beta1, beta2, eps, 1.0f, stream);
}
