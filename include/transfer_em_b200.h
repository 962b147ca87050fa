/* transfer_em_b200 - C ABI of the B200-native transfer_em hot path.
 *
 * The reference (janelia-flyem/transfer_em) has no FFI of its own: its boundary is Python call
 * signatures over Keras objects that delegate the arithmetic to TensorFlow.  This header is the
 * C boundary the replacement puts underneath those signatures; every entry point cites the
 * reference interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 on success, a negative tem_status otherwise; tem_last_error()
 *     returns a thread-local message for the last failure on the calling thread.
 *   - tensors are dense channels-last ([B,Z,Y,X,C]; 2-D data is [B,1,Y,X,C]) and are owned by the
 *     caller (device pointers unless stated).  The library owns only its workspace, parameter /
 *     optimizer arenas and communicator inside a handle.
 *   - every compute call takes a cudaStream_t (passed as void*) and is asynchronous w.r.t. the host
 *     unless its comment says otherwise.
 *   - a handle is not thread-safe; distinct handles are independent.
 *   - there is no CPU fallback: without a CUDA device tem_create fails with TEM_ERR_CUDA.
 */
#ifndef TRANSFER_EM_B200_H
#define TRANSFER_EM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEM_ABI_VERSION 1

typedef enum {
  TEM_OK = 0,
  TEM_ERR_ARG = -1,       /* invalid argument (reference raises RuntimeError: cgan.py:52-53, generator.py:37-38) */
  TEM_ERR_CUDA = -2,      /* CUDA runtime / driver failure */
  TEM_ERR_NCCL = -3,      /* NCCL failure or libnccl not loadable */
  TEM_ERR_STATE = -4,     /* call sequence error (e.g. backward before forward) */
  TEM_ERR_UNSUPPORTED = -5
} tem_status;

typedef enum { TEM_U8 = 0, TEM_BF16 = 1, TEM_F32 = 2 } tem_dtype;
typedef enum { TEM_NET_G = 0, TEM_NET_F = 1, TEM_NET_DX = 2, TEM_NET_DY = 3 } tem_net;
typedef enum { TEM_LOSS_FOCAL = 0, TEM_LOSS_LSGAN_L1 = 1 } tem_loss_mode;

typedef struct tem_handle tem_handle;

/* Mirrors EM2EM.__init__(dimsize, exp_name, is3d, norm_type, ckpt_restore, wf, focal_gamma, disc_prior)
 * (transfer_em/cgan.py:40) plus the optimizer constants of cgan.py:69-73. */
typedef struct {
  int32_t abi_version;     /* TEM_ABI_VERSION */
  int32_t device;          /* CUDA device ordinal */
  int32_t is3d;            /* 1 = 3-D volumes, 0 = 2-D images */
  int32_t wf;              /* width factor: 1,2,4,8,16,32 (cgan.py:49) */
  int32_t dimsize;         /* generator input edge; >= 74 and = 2 (mod 4) (cgan.py:52, generator.py:18) */
  int32_t max_batch;       /* largest per-call batch the workspace is sized for */
  int32_t loss_mode;       /* tem_loss_mode; focal = reference behaviour (cgan.py:78-81) */
  int32_t dropout;         /* 1 = Dropout(0.5) live in training passes (models/utils.py:134) */
  float focal_gamma;       /* cgan.py:40 focal_gamma=2 */
  float lr, beta1, beta2, eps; /* Adam(2e-4, beta_1=0.5), Keras defaults .999 / 1e-7 (cgan.py:69-73) */
  uint64_t seed;           /* weight init N(0,0.02) (models/utils.py:58) and dropout streams */
  int32_t train;           /* 1 = allocate training workspace (6 G + 4 D passes); 0 = inference only */
  int32_t use_tensor_cores;/* 1 = tcgen05 implicit-GEMM kernels where shapes allow; 0 = direct kernels only */
} tem_config;

const char* tem_last_error(void);
int tem_abi_version(void);
void tem_default_config(tem_config* cfg);

/* EM2EM.__init__ (cgan.py:40-103) without the TF checkpoint manager. */
int tem_create(const tem_config* cfg, tem_handle** out);
int tem_destroy(tem_handle* h);

/* model.outdimsize / model.buffer (cgan.py:64-66). */
int tem_out_dim(const tem_handle* h, int32_t* outdimsize, int32_t* buffer);

/* ---- parameters: model.trainable_variables (Keras kernel layouts, layer order g0..g11 / d0..d8,bias) ---- */
int64_t tem_param_count(const tem_handle* h, int net);        /* <0 on error */
int tem_num_variables(const tem_handle* h, int net);
/* shape[0..ndim) in Keras order ([k..,Cin,Cout]; convT [k..,Cout,Cin]); offset in elements into the net's flat vector */
int tem_variable_info(const tem_handle* h, int net, int var, char name[32], int64_t* offset, int32_t* ndim, int64_t shape[6]);
/* which: 0 = parameters, 1 = gradients of the last step, 2 = Adam m, 3 = Adam v.  ptr may be host or device. */
int tem_get_vector(tem_handle* h, int net, int which, float* dst, void* stream);
int tem_set_vector(tem_handle* h, int net, int which, const float* src, void* stream);
int tem_get_step(const tem_handle* h, int64_t* step);
int tem_set_step(tem_handle* h, int64_t step);

/* ---- forward passes: generator_g(x, training=..) / discriminator(x) (generator.py:22, discriminator.py:14) ----
 * in: [B,n,n,n,1] (or [B,1,n,n,1]) of in_dtype.  TEM_U8 input is standardised on load with
 * meanstd = {mean,std}: (u/127.5 - 1 - mean)/std (datasets.py:157-163,193-202); TEM_F32 input is taken as is.
 * out: fp32 [B,n-34,..,1].  dropout_key != 0 enables Dropout(0.5) with that counter-hash key. */
int tem_gen_forward(tem_handle* h, int net, const void* in, int in_dtype, const float* meanstd,
                    int B, int n, uint32_t dropout_key, float* out, void* stream);
/* in: fp32 [B,m,m,m,1]; logits: fp32 [B,l,l,l,1] (l = 1 for m = 40 in 3-D; 6x6 in 2-D). */
int tem_disc_forward(tem_handle* h, int net, const float* in, int B, int m, float* logits, void* stream);
int tem_disc_out_dim(const tem_handle* h, int m, int32_t* l);
/* debug/test access to the stored activation of layer `layer` of the last tem_gen_forward /
 * tem_disc_forward (bf16 as fp32, dense [B,d,d,d,C]); sizes via *count. */
int tem_last_activation(tem_handle* h, int net, int layer, float* dst, int64_t* count, void* stream);

/* ---- EM2EM.train_step(real_x, real_y) (cgan.py:144-230) ----
 * real_x/real_y: [B,n,..,1] of in_dtype (TEM_F32 standardised, or TEM_U8 + meanstd_x/meanstd_y).
 * losses_out: 7 floats in the order of cgan.py:230 (device or pinned-host pointer, written on `stream`).
 * Runs 6 G + 4 D forwards, one combined backward, the gradient all-reduce when a communicator is
 * attached (mean over ranks), and one multi-tensor Adam update for all four networks. */
int tem_train_step(tem_handle* h, const void* real_x, const void* real_y, int in_dtype,
                   const float* meanstd_x, const float* meanstd_y, int B, float* losses_out, void* stream);
/* Same as tem_train_step but stops after the backward pass (no all-reduce, no Adam): gradients are
 * readable with tem_get_vector(which=1).  Used by the parity tests. */
int tem_train_grads(tem_handle* h, const void* real_x, const void* real_y, int in_dtype,
                    const float* meanstd_x, const float* meanstd_y, int B, float* losses_out, void* stream);
/* Apply Adam to the current gradients (after tem_train_grads); grad_scale multiplies the gradients. */
int tem_apply_adam(tem_handle* h, float grad_scale, void* stream);
/* Measurement aid (not a training entry point): captures one train step (forward, backward, Adam; single GPU) into a CUDA
 * graph and replays it `reps` times with the captured dropout keys / learning-rate scalar; writes the mean device time per
 * replay.  tem_last_error() then holds "graph nodes: N".  Replaces nothing in the reference; used by tools/graph_probe.py. */
int tem_debug_graph_replay(tem_handle* h, const void* real_x, const void* real_y, int in_dtype,
                           const float* meanstd_x, const float* meanstd_y, int B, int reps, float* ms_per_step);

/* Output of generator pass p of the last train step (0 fake_y,1 cycled_x,2 fake_x,3 cycled_y,4 same_x,5 same_y). */
int tem_train_output(tem_handle* h, int pass, float* dst, int64_t* count, void* stream);
/* debug/test: gradient w.r.t. the pre-activation of layer `layer` left in the backward scratch by the LAST
 * backward pass of a train step (generator: the same_y pass, cgan.py:181; discriminator: D_x on fake_x). */
int tem_debug_backward_scratch(tem_handle* h, int is_gen, int layer, float* dst, int64_t* count, void* stream);
/* Override the dropout keys of the next train step (6 passes x 2 layers); 0 disables a mask. */
int tem_set_dropout_keys(tem_handle* h, const uint32_t keys[12]);
int tem_get_dropout_keys(tem_handle* h, uint32_t keys[12]);

/* ---- predict_ng_cube (transfer_em/utils.py:41-130) on a device-resident uint8 [VZ,VY,VX] volume ----
 * start/size are (x,y,z) like the reference; out is uint8 [size_z,size_y,size_x] (device).
 * Tiles are the reference's (74^3 at stride 36 for the 74/40 model); tile_z_begin/end select a slab of
 * z tile-layers (0..ceil(size_z/36)) so that ranks can shard the request with no communication.
 * fetch_input != 0 additionally fills in_out (same shape) with the truncated input (utils.py:122-125). */
int tem_predict_volume(tem_handle* h, int net, const uint8_t* vol, const int64_t vol_dims_zyx[3],
                       const int64_t start_xyz[3], const int64_t size_xyz[3],
                       const float meanstd_x[2], const float meanstd_y[2],
                       int outdimsize, int buffer, int tile_z_begin, int tile_z_end,
                       uint8_t* out, uint8_t* in_out, void* stream);

/* ---- data-parallel training (README.md:93-94 / cgan.py:8-11 TODO) ---- */
int tem_comm_unique_id(uint8_t id[128]);
int tem_comm_init(tem_handle* h, const uint8_t id[128], int rank, int world);
int tem_comm_world(const tem_handle* h, int* rank, int* world);
/* broadcast rank 0's parameters + optimizer state to all ranks */
int tem_comm_sync_params(tem_handle* h, void* stream);

/* ---- measurement hooks (bench.py) ---- */
/* number of kernels this library has launched in this process */
uint64_t tem_launch_count(void);
/* name of the kernel function the most recent convolution / weight-gradient call dispatched to (tests assert the path) */
const char* tem_last_kernel(void);
/* bracket every convolution launch with CUDA events on its stream, aggregated by "<layer>.<fwd|dgrad|wgrad>" */
int tem_profile_enable(tem_handle* h, int on);
/* synchronises, then writes one line per tag: "tag count total_ms algorithmic_bytes_per_launch flops_per_launch" */
int tem_profile_report(tem_handle* h, char* buf, int64_t buflen);

/* ---- element-wise conventions, exposed for bit-exact tests ---- */
/* datasets.py:193-202 + 157-163 */
int tem_standardize_u8(const uint8_t* in, float* out, int64_t n, const float meanstd[2], void* stream);
/* utils.py:109,118: (y*std+mean+1)*127.5 -> rint -> wrap to uint8 */
int tem_unstandardize_to_u8(const float* in, uint8_t* out, int64_t n, const float meanstd[2], void* stream);

/* ---- input conditioning on device ---- */
/* datasets.py:123-155 augment(): per sample, output axis k walks input axis perm[k] (tf.transpose), flipped output axes are
   reversed (tf.reverse), then `*= var_adj; += mean_adj` (two fp32 roundings, as in the reference).  in_dtype TEM_U8 fuses
   scale_tensor + standardize_population (meanstd) in front, TEM_F32 takes the already standardised tensor.  All pointers are
   device pointers: in [B, *], out fp32 [B, n0, n1, n2(, 1)], perm / flip int32 [B][3], var_adj / mean_adj fp32 [B]. */
int tem_augment(const void* in, int in_dtype, const float meanstd[2], float* out, int32_t B, const int32_t out_dims[3],
                const int32_t* perm, const int32_t* flip, const float* var_adj, const float* mean_adj, void* stream);
/* datasets.py:173-190 get_meanstd(), one tensor: out[0] = tf.math.reduce_mean, out[1] = tf.math.reduce_variance (population).
   scratch: 32 zeroed bytes of device memory, left zeroed. */
int tem_mean_var(const float* in, int64_t n, void* scratch, float* out, void* stream);
/* debug.py:7-63 warp_tensor() on one [Z,Y,X] (2-D: Z = 1, ndims = 2) fp32 tensor: box blur 3^ndims ('SAME', zero padding,
   filter 1/27 or 1/9), then every voxel whose 4^ndims 'SAME' window (offsets -1..+2) holds a seed uniform[v] < hole_rate
   (the reference draws tf.random.uniform; here the caller passes the draw) is set to the mean of the blurred tensor.
   scratch: 8 bytes of device memory.  in and out must not alias. */
int tem_warp_tensor(const float* in, const float* uniform, float* out, const int32_t dims_zyx[3], int32_t ndims,
                    float hole_rate, void* scratch, void* stream);

/* ---- chunked output ---- */
/* model_cloudrun/transferem.py:171-184: re-tiles the uint8 result volume vol[z,y,x] into chunk^3 blocks (clipped at the volume
   edge), each block contiguous in C order, blocks concatenated z-outer / y / x-inner: block (bz,by,bx) starts at byte
   z0*Y*X + cz*(y0*X + cy*x0) of out (z0 = bz*chunk, cz = its clipped depth, ...).  out has Z*Y*X bytes. */
int tem_chunk_volume(const uint8_t* vol, const int64_t dims_zyx[3], int32_t chunk, uint8_t* out, void* stream);

/* ---- per-op entry points (tests, layer-level parity) ---- */
typedef struct {
  /* geometry */
  int32_t B;
  int32_t in_dims[3];      /* Z,Y,X of the input tensor */
  int32_t cin, cout;
  int32_t k[3];            /* kernel extent per axis */
  int32_t stride[3];
  int32_t transposed;      /* 0 = Conv (VALID), 1 = Conv*DTranspose(k=4,s=2,'same') (models/utils.py:129-130) */
  float slope;             /* LeakyReLU slope (1 = linear) */
  uint32_t dropout_key;    /* 0 = none */
  int32_t in_dtype;        /* tem_dtype of `in` (u8 only with cin == 1) */
  int32_t out_dtype;       /* TEM_BF16 or TEM_F32 */
  float meanstd[2];        /* for u8 input */
  int32_t use_tensor_cores;
} tem_conv_desc;
/* y = act(dropout(conv(x, w) + bias)).  w: fp32 Keras layout.  out dims returned in out_dims. */
int tem_conv_forward(const tem_conv_desc* d, const void* in, const float* w, const float* bias,
                     void* out, int32_t out_dims[3], void* stream);
/* dx = conv_dgrad(dy, w) * lrelu'(x_act) (x_act = stored activation of the producer, may be NULL) */
int tem_conv_dgrad(const tem_conv_desc* d, const void* dy, int dy_dtype, const float* w,
                   const void* x_act, float x_slope, void* dx, int dx_dtype, void* stream);
/* dw (+)= wgrad(x, dy) in Keras layout, fp32 */
int tem_conv_wgrad(const tem_conv_desc* d, const void* x, const void* dy, int dy_dtype, float* dw, void* stream);

/* focal losses (cgan.py:110-142).  loss_out += scale * mean(l); grad = scale * dl/dx / n */
int tem_focal_logits(const float* logits, int64_t n, float target, float gamma, float scale,
                     float* loss_out, float* grad, void* stream);
/* a: reference image (fp32), b: generated (fp32), same dense shape n */
int tem_focal_probs(const float* a, const float* b, int64_t n, float gamma, float scale,
                    float* loss_out, float* grad_b, void* stream);
/* Keras Adam on flat vectors (cgan.py:69-73,218-228); step = 1-based step after increment */
int tem_adam(float* p, const float* g, float* m, float* v, int64_t n, int64_t step,
             float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
/* dropout keep mask of the counter hash (1.0 / 0.0), for tests */
int tem_dropout_mask(uint32_t key, float* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRANSFER_EM_B200_H */
