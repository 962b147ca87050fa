// Host-side runtime structures of libtem_b200: network tables, workspaces, pass executors.
#pragma once
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "../../include/transfer_em_b200.h"
#include "tem_kernels.cuh"

struct LayerSpec {
  char name[8];
  int transposed;     // 0 conv, 1 convT
  int k, stride;
  int cin, cout;
  float slope;        // activation slope (1 = linear)
  int dropout, bias;
  long long w_off, b_off;   // element offsets inside the net's flat parameter vector
  long long w_count;
};

struct NetSpec {
  int is_gen;
  std::vector<LayerSpec> L;
  long long count;        // parameters in this net
  long long arena_off;    // offset of this net in the shared arena
};

struct Tensor {          // dense [B, d0, d1, d2, C]
  void* p; int dtype; int d[3]; int C;
  long long per_sample() const { return (long long)d[0] * d[1] * d[2] * C; }
};

struct InputRef {        // first-layer input of a pass
  const void* p; int dtype;
  int dims[3];           // tensor dims
  int shift[3];          // logical -> tensor coordinate
  const int* origins;    // per-sample origins (tiled inference)
  int use_lut; float mean, stdv;
};

struct GenPass {
  Tensor a[12];          // a[0..10] bf16 activations, a[11] fp32 output
  InputRef in;
  uint32_t keys[2];      // dropout keys of g6, g9 (0 = off)
  int B, n;
  bool valid;
};

struct DiscPass {
  Tensor a[9];           // a[0..7] bf16, a[8] fp32 logits
  InputRef in;
  int B, m;
  bool valid;
};

struct NcclApi;

// optional per-launch CUDA-event timing, aggregated by (layer, op) tag: used by bench.py for the roofline line
struct ProfRec { char tag[24]; const char* kernel; cudaEvent_t e0, e1; double bytes, flops; };
extern const char* g_tem_last_kernel;   // name of the kernel the last dispatch chose (profiling only)
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
};

struct tem_handle {
  tem_config cfg;
  NetSpec nets[4];
  long long total_params;      // all four nets
  long long arena_elems;       // total_params + 16 loss slots (all-reduced together)
  float *params, *grads, *adam_m, *adam_v;
  float* loss_dev;             // = grads + total_params  (16 floats)
  int64_t step;
  int nd;                      // 3 or 2
  int n, outdim, buffer;       // generator geometry
  int dm, dl;                  // discriminator input / logit edge
  int maxB;
  // workspaces
  GenPass gp[7];               // 0..5 train passes, 6 inference / API pass
  DiscPass dp[5];              // 0..3 train passes (DxR, DyR, DxF, DyF), 4 API pass
  Tensor gdP[4][11];           // generator backward scratch (bf16): one set per internal stream
  Tensor ddP[4][8];            // discriminator backward scratch
  float* dOut[6];              // gradient w.r.t. each generator pass output (fp32)
  float* dlog[6];              // logit gradients: gen_y, gen_x, dyr, dyf, dxr, dxf
  int* tile_origins; int* tile_index;   // device, maxB*3 each
  int* h_tile_origins; int* h_tile_index;   // pinned host staging
  long long tile_cap;                       // tiles the device / host tile tables hold (grown per request)
  uint32_t next_keys[12]; bool keys_overridden;
  // communicator
  void* comm; int rank, world;
  std::vector<void*> allocs;
  int last_gen_net, last_disc_net;
  Profiler prof;
  // packed bf16 UMMA weight images, keyed by (weight pointer, form, columns); re-packed when params change
  struct Packed { bf16* buf; size_t bytes; uint64_t version; ConvArgs args; int kind; };   // kind: 0 conv_tc3, 2 conv_tc_s2, 3 conv_tcw
  std::map<std::tuple<const float*, int, int>, Packed> packed;
  uint64_t params_version;
  // four internal streams overlap the independent passes of a train step (G(real_x) || F(real_y), ...)
  cudaStream_t aux[6];         // A, B: the chained generator passes (high priority: the critical path); C, D: identity passes; E, F: discriminator passes
  cudaEvent_t ev[16];
  bool overlap_ready;
  bool in_overlap;              // a train step is between its stream fork and join
};

struct ProfScope {
  tem_handle* h; cudaStream_t st; size_t idx; bool live;
  ProfScope(const tem_handle* hc, const char* layer, const char* op, double bytes, double flops, cudaStream_t s);
  ~ProfScope();
};
extern unsigned long long g_tem_launches;

void tem_set_error(const char* fmt, ...);
#define TEM_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      tem_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), cudaGetErrorString(_e)); \
      return TEM_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)
#define TEM_CHECK(expr)                    \
  do {                                     \
    int _s = (expr);                       \
    if (_s != TEM_OK) return _s;           \
  } while (0)
