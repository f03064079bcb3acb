/*
 * dsg_b200.h — C ABI of libdsg_b200.so, the sm_100a denoising engine behind the `diffusers` call surface
 * used by SS47816/DriveSceneGen.
 *
 * The reference has NO native FFI for this path: the boundary it binds is the Python import surface of
 * diffusers==0.20.0 / accelerate==0.22.0 (requirements.txt:15-16).  Each entry point below names the upstream
 * operation it replaces and the reference call site that reaches it (paths relative to /root/reference):
 *   UNet2DModel.forward     DriveSceneGen/pipeline/training_pipeline.py:84, DriveSceneGen/scripts/generation.py:14
 *   DDPMScheduler.step      DriveSceneGen/pipeline/training_pipeline.py:26-32, DriveSceneGen/scripts/generation.py:14-20
 *   DDPMScheduler.add_noise DriveSceneGen/pipeline/training_pipeline.py:80
 *   pipeline post-process   DriveSceneGen/scripts/generation.py:22-24 ((x/2+.5).clamp(0,1) -> NHWC -> uint8)
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; dsg_last_error() gives a thread-local message.
 *   - all pointers are DEVICE pointers unless the name ends in _host; the library never allocates, frees or
 *     retains caller memory, never synchronises the host, and launches on the given stream, so every call is
 *     legal under CUDA-graph stream capture.
 *   - activations between U-Net ops are NHWC fp16 ("h16"): [N][H][W][C], C a multiple of 64, base pointers
 *     16-byte aligned.  Model input / output and the schedulers work on NCHW fp32 like the reference.
 *   - stream is a cudaStream_t passed as void*.
 */
#ifndef DSG_B200_H_
#define DSG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSG_OK 0
#define DSG_ERR_INVALID (-1)
#define DSG_ERR_CUDA (-2)
#define DSG_ERR_UNSUPPORTED (-3)

int dsg_version(void);
const char* dsg_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches evidence) */
int64_t dsg_launch_count(void);
/* a host that replays a CUDA graph captured from this library's calls reports the graph's kernel count here (the
 * replay itself does not pass through the library) */
void dsg_count_graph_launches(int64_t n);
/* 1 if the current device is sm_100 (tcgen05/TMA kernels can run), 0 otherwise, <0 on error */
int dsg_device_ok(void);

/* ---------------------------------------------------------------- schedulers (fp32, NCHW, elementwise) ----
 * coef_table: device float[rows][8]; row r =
 *   { sqrt(1-abar_t), sqrt(abar_t), c_x0, c_xt, sigma, clip_range, has_noise(0/1), unused }     (DDPM)
 *   { sqrt(1-abar_t), sqrt(abar_t), sqrt(abar_prev), dir_coef, sigma, clip_range, has_noise, unused } (DDIM)
 * The row is row_dev ? *row_dev : row (a device-side index keeps one captured graph valid for every step).
 * Arithmetic order follows upstream scheduling_ddpm.py::step / scheduling_ddim.py::step exactly, with every
 * product and sum rounded separately (no FMA contraction), so results are bit-identical to the fp32 CPU path. */
int dsg_ddpm_step(const float* eps, const float* sample, const float* noise /* may be NULL */, float* prev,
                  int64_t numel, const float* coef_table, const int32_t* row_dev, int32_t row, void* stream);
int dsg_ddim_step(const float* eps, const float* sample, const float* noise /* may be NULL */, float* prev,
                  int64_t numel, const float* coef_table, const int32_t* row_dev, int32_t row, void* stream);
/* DDPMScheduler.add_noise: out = sqrt_ac[t[n]] * x0 + sqrt_1mac[t[n]] * noise ; t is int64[batch].  The tables hold
 * table_len entries; a timestep outside [0, table_len) (IndexError upstream) yields an all-NaN sample, never an
 * out-of-bounds read. */
int dsg_add_noise(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac,
                  const float* sqrt_1mac, int32_t table_len, float* out, int32_t batch, int64_t per_sample,
                  void* stream);
/* Sampling-loop bookkeeping of DDPMPipeline.__call__ (`for t in self.scheduler.timesteps`, reached from
 * DriveSceneGen/scripts/generation.py:14) on the device: state = {k, n_steps}; t = schedule[min(k, n_steps-1)];
 * t_f[0..batch) = t; *row = t; state[0] = k + 1.  First node of the captured step graph: a replay then takes no
 * host-side argument updates. */
int dsg_step_advance(const int32_t* schedule, int32_t* state, float* t_f, int32_t batch, int32_t* row, void* stream);
/* pipeline post-process: NCHW fp32 latent -> NHWC; u8 = round(clamp(x/2+.5,0,1)*255) (numpy_to_pil) and/or
 * f32 = clamp(x/2+.5,0,1).  Either output may be NULL. */
int dsg_latent_to_image(const float* latent, uint8_t* out_u8, float* out_f32, int32_t n, int32_t c, int32_t h,
                        int32_t w, void* stream);

/* ---------------------------------------------------------------- raster images either side of the path ---- */
/* Input side — replaces the per-item arithmetic of Image_Dataset.__getitem__
 * (DriveSceneGen/utils/datasets/dataset.py:20-23,44-47): ToTensor (x / 255) then Normalize([0.5], [0.5])
 * ((x - 0.5) / 0.5), each fp32 operation rounded separately like torchvision.  img: uint8 [n][h][w][c_img] (what PIL
 * decodes), out: float [n][c_out][h][w] = channels 0..c_out-1.  The Resize in the reference transform is the identity
 * when the stored raster already has the model's size; other sizes are the caller's business. */
int dsg_image_to_sample(const uint8_t* img, float* out, int32_t n, int32_t h, int32_t w, int32_t c_img, int32_t c_out,
                        void* stream);
/* The same with the transform's Resize((out_h, out_w), antialias=False) (dataset.py:20-23) between ToTensor and
 * Normalize — the reference rasterises at 512^2 (config/data_rasterization.yaml:6) and trains at 256^2
 * (scripts/train.py:14-15), so this resample is live on every sample.  Bit-exact restatement of ATen's CPU
 * upsample_bilinear2d (align_corners = False); mode selects which of ATen's two CPU kernels is reproduced:
 * 0 = the generic N-d kernel (multi-threaded hosts, out_h + out_w > 128: the reference's case), 1 = the channels-last
 * kernel (single-threaded hosts with 3 channels, or out_h + out_w <= 128).  img: uint8 [n][h][w][c_img], or
 * float [n][h][w][c_img] already in [0, 1] when img_is_f32 (the `.pkl` branch, dataset.py:38-42, skips ToTensor). */
int dsg_resize_to_sample(const void* img, int32_t img_is_f32, float* out, int32_t n, int32_t h, int32_t w, int32_t c_img,
                         int32_t c_out, int32_t out_h, int32_t out_w, int32_t mode, void* stream);
/* Vectoriser front end — replaces get_gray_image
 * (DriveSceneGen/vectorization/utils/image_utils.py:13-42) for a batch of rasters that are still on the device.
 *   img   uint8 [n][h][w][c], c = 3 or 4 (channel 0 = dx, 1 = dy, 2 = speed)
 *   hist  uint32 [n][3][256] out: np.histogram(channel / 255.0, bins=256, range=(0, 1)) (zeroed by the call)
 *   peaks int32 [n][3] out: np.argmax of each histogram (first maximum); the peak value is peaks / 256.0
 *   mask  uint8 [n][h][w] out: 0 where |dx/255 - peak_dx/256| <= thresh and |dy/255 - peak_dy/256| <= thresh (float64
 *         comparisons like combine_dx_dy, image_utils.py:6-10), else 255
 *   gray3 uint8 [n][h][w][3] out, may be NULL: the mask repeated over three channels (what Image.fromarray receives) */
int dsg_gray_mask(const uint8_t* img, uint32_t* hist, int32_t* peaks, uint8_t* mask, uint8_t* gray3, int32_t n,
                  int32_t h, int32_t w, int32_t c, double thresh, void* stream);
/* Agent blobs — replaces the per-pixel head of extract_agents
 * (DriveSceneGen/vectorization/direct/extract_vehicles.py:136-148): img = (x * 255).astype(uint8) (truncation, x in
 * [0, 1]), cv2.cvtColor of three equal channels (identity), cv2.threshold(gray, thresh, 255, THRESH_BINARY).
 * plane: n float planes of hw values (the speed channel), plane i at plane + i * plane_stride; out uint8 [n][hw].
 * cv2.findContours / minAreaRect stay on the host. */
int dsg_agent_threshold(const float* plane, int64_t plane_stride, uint8_t* out, int32_t n, int64_t hw, int32_t thresh,
                        void* stream);

/* ---------------------------------------------------------------- U-Net building blocks ------------------- */
/* Timesteps + TimestepEmbedding + all per-ResnetBlock time_emb_proj in two launches.
 *   t: float[batch] (timestep values), freqs: float[half] (exp table, host-computed like upstream),
 *   w1t [2*half][hidden], w2t [hidden][hidden]: the two TimestepEmbedding weights TRANSPOSED ([in][out]), b1, b2,
 *   wp [proj_total][hidden], bp [proj_total]  (all fp32)
 *   emb_ws: float[batch][hidden] scratch; out: float[batch][proj_total] = Linear(SiLU(emb)) per block. */
int dsg_time_embed(const float* t, const float* freqs, int32_t half, int32_t flip_sin_to_cos, const float* w1t,
                   const float* b1, const float* w2t, const float* b2, int32_t hidden, const float* wp,
                   const float* bp, int32_t proj_total, float* emb_ws, float* out, int32_t batch, void* stream);

/* conv_in: NCHW fp32 [n][cin][h][w] (cin <= 4) -> h16 [n][h][w][cout]; 3x3, pad 1.  w: fp32 [cout][cin][3][3]. */
int dsg_conv_in(const float* x, const float* w, const float* b, void* out_h16, int32_t n, int32_t cin, int32_t h,
                int32_t wd, int32_t cout, void* stream);
/* dsg_conv_in on (*xscale) * x, xscale a device pointer to a power of two: conv_out's data gradient, whose fp32 input
 * (the raw loss gradient) is far below fp16 range.  cout == 64, cin <= 3: tensor cores (x * xscale rounded to fp16);
 * other shapes: the fp32 CUDA-core kernel with the scale folded into the weights. */
int dsg_conv_in_scaled(const float* x, const float* xscale, const float* w, const float* b, void* out_h16, int32_t n,
                       int32_t cin, int32_t h, int32_t wd, int32_t cout, void* stream);
/* conv_in + the per-channel GroupNorm totals of its output (same int64 fixed-point format as dsg_gn_stats; added to
 * `stats`, which the caller zeroes): cout == 64 and cin <= 3 run on the tensor cores (mma.sync, fp16 operands,
 * fp32 accumulate) with the statistics fused; any other shape = dsg_conv_in followed by dsg_gn_stats. */
int dsg_conv_in_stats(const float* x, const float* w, const float* b, void* out_h16, void* stats, int32_t n,
                      int32_t cin, int32_t h, int32_t w_, int32_t cout, void* stream);
/* conv_out: h16 [n][h][w][cin] (already GroupNorm+SiLU'd) -> NCHW fp32 [n][cout][h][w] (cout <= 4). */
int dsg_conv_out(const void* x_h16, const float* w, const float* b, float* out, int32_t n, int32_t cin, int32_t h,
                 int32_t wd, int32_t cout, void* stream);

/* conv_norm_out + SiLU + conv_out in one pass (inference): x_h16 is the RAW 64-channel tensor, gn_coef the float2
 * [n][64] coefficients of dsg_gn_coef ((a/2, b/2) with GroupNorm(x) = a x + b; NULL = x is already activated).
 * cin == 64, cout <= 8; warp-level mma.sync (fp16 operands, fp32 accumulate), NCHW fp32 output. */
int dsg_conv_out_fused(const void* x_h16, const float* gn_coef, const float* w, const float* b, float* out, int32_t n,
                       int32_t cin, int32_t h, int32_t wd, int32_t cout, void* stream);

/* GroupNorm over the channel-concatenation of up to two h16 tensors (x1 has c1 channels, x2 has c2, x2 may be NULL).
 * Statistics are per-channel fixed-point totals: stats int64 [n][c][2] = { sum(x) * 2^24, sum(x^2) * 2^20 },
 * accumulated with integer atomics (order-independent, bit-reproducible).  They come either from dsg_conv's
 * epilogue (out_stats) or from dsg_gn_stats, which ADDS one tensor's totals to a caller-zeroed buffer. */
int dsg_gn_stats(const void* x, int32_t c, void* stats, int32_t n, int64_t hw, void* stream);
/* y = act(GroupNorm(cat(x1,x2))) written as ONE h16 tensor with c1+c2 channels; act: 0 none, 1 SiLU. */
int dsg_gn_apply(const void* x1, int32_t c1, const void* stats1, const void* x2, int32_t c2, const void* stats2,
                 const float* gamma, const float* beta, float eps, int32_t act, void* y, int32_t n, int64_t hw,
                 int32_t groups, void* stream);

/* Implicit-GEMM convolution on tcgen05 tensor cores (TMA -> smem -> UMMA -> TMEM -> epilogue).
 * mode: 0 = 3x3 stride 1 pad 1, 1 = 3x3 stride 2 pad 1, 2 = nearest-2x upsample followed by 3x3 pad 1
 *       (evaluated as four 2x2 sub-pixel convolutions on pre-summed weights), 3 = 1x1,
 *       4 = the adjoint of mode 2 (4x4 stride-2 gather over the four parity views of x; output h/2 x w/2) — used
 *       for the data gradient of the upsample conv with weights packed in mode 12.
 * x: h16 [n][h][w][cin].  sc1/sc2: optional extra 1x1 ("conv_shortcut") inputs at the OUTPUT resolution whose
 * channels are appended to the GEMM K dimension (sc2 may be NULL; csc2 = 0).  residual: optional h16 tensor shaped
 * like the output, added in the epilogue.  bias: float[cout].  temb: optional float[n][temb_stride], element
 * [n][temb_off + c] is added to output channel c.  wpacked: h16, layout produced by dsg_pack_conv_weight.
 * out: h16 [n][oh][ow][cout]. */
typedef struct dsg_conv_args {
  int32_t mode;
  int32_t n, h, w, cin, cout;
  const void* x;
  const void* sc1;
  int32_t csc1;
  const void* sc2;
  int32_t csc2;
  const void* wpacked;
  const float* bias;
  const float* temb;
  int32_t temb_stride, temb_off;
  const void* residual;
  void* out;
  int32_t block_n; /* 0 = auto, else 64/128/256 */
  int32_t impl;    /* 0 = tcgen05 path (CTA-pair halo-reuse kernel where it applies, else the single-CTA halo-reuse
                      kernel, else the tap-streaming kernel),
                      1 = plain CUDA-core cross-check kernel (slow; debugging/tests only),
                      2 = force the tap-streaming tcgen05 kernel, 3 = force the single-CTA halo-reuse tcgen05 kernel,
                      4 = force the CTA-pair (cta_group::2) halo-reuse kernel
                      (3/4: DSG_ERR_UNSUPPORTED when the mode/shape is outside it: modes 0 and 2, W >= 8, H >= tile) */
  /* conv_out form (replaces UNet2DModel.conv_out, 64 -> 3 channels): when out_nchw_f32 is non-NULL the weights are
   * packed for cout = 16 (rows >= cout_real are zero), mode must be 0, `out` is ignored and the first cout_real
   * output channels are written as NCHW fp32 [n][cout_real][h][w].  Halo-reuse kernel only (W >= 8, H >= 18). */
  void* out_nchw_f32;
  int32_t cout_real;
  /* optional: per-channel GroupNorm totals of the OUTPUT, int64 [n][cout][2] = { sum * 2^24, sum of squares * 2^20 },
   * ADDED to (zero the buffer first) from the epilogue — the statistics pass of the next GroupNorm for free.
   * The tcgen05 kernels store the total of channels (2k, 2k+1) in channel 2k's slot and leave 2k+1's untouched:
   * valid for dsg_gn_apply whenever the group size and c1 are even (it only sums whole groups). */
  void* out_stats;
  /* optional fused GroupNorm + SiLU of the INPUT (modes 0 only, halo-reuse kernels — check dsg_conv_gn_fusable):
   * when gn_coef is non-NULL, x (and x2) hold the RAW tensor(s) the GroupNorm reads — x has cin1 channels, x2 (may
   * be NULL, then cin1 = cin) the remaining cin - cin1, i.e. the up-block torch.cat is fused too — and the conv input
   * is SiLU(GroupNorm(cat(x, x2))) formed in shared memory from gn_coef = float[n][cin][2] (dsg_gn_coef). */
  const void* x2;
  int32_t cin1;
  const float* gn_coef;
} dsg_conv_args;
int dsg_conv(const dsg_conv_args* args, void* stream);
/* 1 if dsg_conv can run this shape with gn_coef (fused GroupNorm + SiLU input), else 0 */
int dsg_conv_gn_fusable(int32_t mode, int32_t h, int32_t w, int32_t cin, int32_t cin1, int32_t cout);
/* Per-(sample, channel) GroupNorm coefficients for the fused form: coef[n][c] = { a / 2, b / 2 } with
 * GroupNorm(x) = a x + b (a = gamma * rstd, b = beta - mean * a), from the same statistics dsg_gn_apply takes. */
int dsg_gn_coef(int32_t c1, const void* stats1, int32_t c2, const void* stats2, const float* gamma, const float* beta,
                float eps, float* coef, int32_t n, int64_t hw, int32_t groups, void* stream);
/* K-extent (in fp16 elements per output channel row) and row count of the packed weight for a mode. */
int64_t dsg_packed_k(int32_t mode, int32_t cin, int32_t csc);
int64_t dsg_packed_rows(int32_t mode, int32_t cout);
/* All conv weights of a model in one launch, staged through shared memory (pack_weights.cu; bit-identical to
 * dsg_pack_conv_weight).  jobs_dev: DEVICE array of njobs entries, each describing one dsg_pack_conv_weight call (same
 * field meaning), sorted by chunk_begin = the running sum of dsg_pack_job_blocks(mode, cout, cin, csc) over the
 * preceding jobs; total_blocks = that sum over all jobs.  dsg_pack_job_blocks returns -1 for a layer too wide for the
 * staged form (its source row exceeds 40 KB): pack that one with dsg_pack_conv_weight. */
typedef struct dsg_pack_job {
  int32_t mode, cout, cin, csc;
  const float* w;
  const float* w_sc;
  void* out;
  int64_t k_total, rows, chunk_begin;
} dsg_pack_job;
int64_t dsg_pack_job_blocks(int32_t mode, int32_t cout, int32_t cin, int32_t csc);
int dsg_pack_conv_weights_batched(const dsg_pack_job* jobs_dev, int32_t njobs, int64_t total_blocks, void* stream);
/* Data-gradient ("dgrad") packings: dsg_pack_conv_weight modes 10..13 take the SAME fp32 OIHW weight as modes 0..3
 * (cout / cin are the forward conv's) and produce the weights of the conv that maps the OUTPUT gradient to the INPUT
 * gradient: 10 -> run dsg_conv mode 0 (cin' = cout, cout' = cin), 11 -> mode 2 over the low-resolution gradient,
 * 12 -> mode 4 over the high-resolution gradient, 13 -> mode 3.  K / rows of those packings: */
int64_t dsg_packed_k_dgrad(int32_t fwd_mode, int32_t cout);
int64_t dsg_packed_rows_dgrad(int32_t fwd_mode, int32_t cin);
/* Pack fp32 OIHW conv weights (+ optional 1x1 shortcut weights [cout][csc]) into the h16 GEMM layout.
 * All pointers are device pointers. */
int dsg_pack_conv_weight(int32_t mode, const float* w_oihw, int32_t cout, int32_t cin, const float* w_sc,
                         int32_t csc, void* wpacked, void* stream);

/* Multi-head self-attention core on tokens: qkv h16 [n][tokens][3*c] (q | k | v, heads x head_dim inside each),
 * out h16 [n][tokens][c]; softmax(q k^T / sqrt(head_dim)) v per head.  head_dim must be 8, 16, 32 or 64. */
int dsg_attention(const void* qkv, void* out, int32_t n, int32_t tokens, int32_t heads, int32_t head_dim,
                  void* stream);
/* Same with an explicit implementation: impl 0 = auto (tcgen05 kernel for head_dim 8 and tokens a multiple of 128 up
 * to 4096, else the CUDA-core flash kernel), 1 = CUDA-core kernel, 2 = tcgen05 kernel or DSG_ERR_UNSUPPORTED.
 * dbg (tests; may be NULL): float[128*128 + 128*16] receives the raw scores of the first 128 x 128 block and the
 * un-normalised output tile (column 8 = softmax denominator) of the first (sample, head). */
int dsg_attention_ex(const void* qkv, void* out, int32_t n, int32_t tokens, int32_t heads, int32_t head_dim,
                     int32_t impl, float* dbg, void* stream);


/* ================================================================ training path (backward + optimizer) ==========
 * Replaces what torch autograd / cuDNN / torch.optim run behind `accelerator.backward(loss)`,
 * `accelerator.clip_grad_norm_` and `optimizer.step()` (DriveSceneGen/pipeline/training_pipeline.py:86-89) for the
 * UNet2DModel of DriveSceneGen/scripts/train.py:39-57.  Activation gradients are h16 like the activations and are
 * carried times a power-of-two factor s (dsg_grad_scale); every parameter-gradient entry point takes `inv_scale`, a
 * DEVICE pointer to 1/s (NULL = 1), and writes fp32 gradients in the parameter's own (torch) layout. */

/* scale[0] = s, scale[1] = 1/s with amax(|dout|) * s in [1, 2) (s = 1 when amax is 0 or not finite).
 * partial: float[parts] scratch. */
int dsg_grad_scale(const float* dout, int64_t numel, float* partial, int32_t parts, float* scale, void* stream);

/* same as dsg_time_embed; saved (may be NULL): float[batch * (2*half + 3*hidden)] = four dense arrays for the
 * backward: sinusoid [batch][2*half], then [batch][hidden] each of the pre-activation of linear_1, SiLU of it, and
 * the pre-activation of linear_2. */
int dsg_time_embed_ex(const float* t, const float* freqs, int32_t half, int32_t flip_sin_to_cos, const float* w1t,
                      const float* b1, const float* w2t, const float* b2, int32_t hidden, const float* wp,
                      const float* bp, int32_t proj_total, float* emb_ws, float* out, int32_t batch, float* saved,
                      void* stream);
/* small-batch fp32 linear backward (time-embedding path).  w is [rows][cols] = torch [out][in].
 *   dgrad: dx[n][k] = (sum_r dy[n][dy_off + r] * w[r][k]) * (pre ? SiLU'(pre[n][k]) : 1)
 *   wgrad: dw[r][k] = inv_scale * sum_n dy[n][dy_off + r] * x[n][k];  db[r] = inv_scale * sum_n dy[n][dy_off + r] */
int dsg_lin_dgrad_small(const float* dy, int32_t ldy, int32_t dy_off, const float* w, int32_t rows, int32_t cols,
                        const float* pre, float* dx, int32_t batch, void* stream);
int dsg_lin_wgrad_small(const float* dy, int32_t ldy, int32_t dy_off, const float* x, int32_t ldx, int32_t rows,
                        int32_t cols, int32_t batch, const float* inv_scale, float* dw, float* db /* may be NULL */,
                        void* stream);

/* GroupNorm(+SiLU) backward over cat(x1, x2) (see dsg_gn_apply for the forward and the statistics format).
 *   dy: h16 [n][hw][c1+c2] gradient w.r.t. the forward OUTPUT.  partial: scratch of
 *   n * (chunks + 1) * (c1+c2) * 2 + n * (c1+c2) * 4 floats: [n][chunks + 1][c1+c2][2] (chunk partials, then one slot per
 *   sample holding the sums behind d beta / d gamma, read by dsg_gn_bwd_params) followed by the per-channel coefficients
 *   of dx [n][c1+c2][4] that a small kernel between the two passes derives from them; 1 <= chunks <= 64.
 *   addend (may be NULL): h16 [n][hw][c1+c2] added to the input gradient (the ResnetBlock shortcut's gradient).
 *   dx1 / dx2: h16 gradients of x1 / x2; accN != 0 adds to the existing content (tensor with two consumers).
 *   parts: CTAs per sample of the apply pass (0 = automatic, only without column sums).  Optional per-CTA column sums
 *   (finish them with dsg_colsum_finalize):
 *     colsum float[n][parts][c1+c2]: of the GroupNorm term of dx (before addend / accumulate) — the time-embedding /
 *            conv1-bias gradient of a ResnetBlock;
 *     osum1 float[n][parts][c1], osum2 float[n][parts][c2]: of the FINAL values stored to dx1 / dx2 — when this call is
 *            the last writer of that gradient tensor, its producer's bias gradient without another read. */
int dsg_gn_bwd(const void* dy, const void* x1, int32_t c1, const void* stats1, const void* x2, int32_t c2,
               const void* stats2, const float* gamma, const float* beta, float eps, int32_t act, float* partial,
               int32_t chunks, const void* addend, void* dx1, int32_t acc1, void* dx2, int32_t acc2, float* colsum,
               float* osum1, float* osum2, int32_t parts, int32_t n, int64_t hw, int32_t groups, void* stream);
/* d gamma[c] / d beta[c] = inv_scale * sum over samples of the per-sample slot dsg_gn_bwd left in `partial` */
int dsg_gn_bwd_params(const float* partial, int32_t n, int32_t chunks, int32_t c, const float* inv_scale,
                      float* dgamma, float* dbeta, void* stream);
/* column sums of an h16 [rows][c] tensor: partial float[parts][c] */
int dsg_colsum_h16(const void* x, int64_t rows, int32_t c, float* partial, int32_t parts, void* stream);
/* partial float[n][parts][c] -> per_n[i][per_n_off + c] (raw per-sample sums, may be NULL) and
 * total[c] = total2[c] = inv_scale * sum over samples and parts (either may be NULL) */
int dsg_colsum_finalize(const float* partial, int32_t n, int32_t parts, int32_t c, float* per_n, int32_t per_n_stride,
                        int32_t per_n_off, const float* inv_scale, float* total, float* total2, void* stream);

/* Weight gradient of a dsg_conv-style convolution on tcgen05 tensor cores (pixel axis = GEMM K, split over CTAs,
 * deterministic two-stage reduction).  mode = the FORWARD conv's mode (0, 1, 2, 3); n, h, w, cin describe the forward
 * input x (h16 [n][h][w][cin]); dy is the h16 output gradient ([n][h][w][cout], [n][h/2][w/2][cout] for mode 1,
 * [n][2h][2w][cout] for mode 2).  grad: fp32 OIHW [cout][ci_total][3][3] ([cout][ci_total] for mode 3); this call
 * fills input-channel columns [ci_off, ci_off + cin) — a conv_shortcut over cat(x1, x2) takes one call per source.
 * workspace: >= dsg_wgrad_workspace_bytes(...) bytes, 16-byte aligned.  impl: 0 = auto (tcgen05 when W >= 8 and
 * H >= pixel tile + 2, else the CUDA-core kernel), 1 = CUDA-core cross-check kernel, 2 = tcgen05 or error. */
typedef struct dsg_wgrad_args {
  int32_t mode;
  int32_t n, h, w, cin, cout;
  const void* x;
  const void* dy;
  float* grad;
  int32_t ci_total, ci_off;
  int32_t accumulate; /* 0 = overwrite, 1 = add to grad */
  const float* inv_scale;
  void* workspace;
  int64_t workspace_bytes;
  int32_t impl;
} dsg_wgrad_args;
int dsg_conv_wgrad(const dsg_wgrad_args* args, void* stream);
int64_t dsg_wgrad_workspace_bytes(int32_t mode, int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout);

/* conv_out data gradient = dsg_conv_in over the fp32 NCHW output gradient with wt = s * flipped, transposed
 * conv_out weight: w fp32 [cout][cin][3][3] -> wt fp32 [cin][cout][3][3]; scale = device pointer to s (NULL = 1). */
int dsg_conv_out_dgrad_weight(const float* w, int32_t cout, int32_t cin, const float* scale, float* wt, void* stream);
/* Weight gradient of conv_in / conv_out: wide = the h16 [n][h][w][wc] side, narrow = the fp32 NCHW [n][nc][h][w] side.
 *   conv_out_form = 0 (conv_in):  dw fp32 [wc][nc][3][3] = inv_scale * sum wide[q][wc] * narrow[c][q + tap]
 *   conv_out_form = 1 (conv_out): dw fp32 [nc][wc][3][3] = inv_scale * sum wide[q + tap][wc] * narrow[c][q]
 * narrow_sum (may be NULL): float[nc] inv_scale * plain sums of the narrow tensor (conv_out's bias gradient).
 * partial: float[parts + 1][nc*9*wc + nc] scratch; parts = number of CTAs (grid-stride over image rows). */
int dsg_small_wgrad(const void* wide_h16, const float* narrow_nchw, int32_t n, int32_t h, int32_t w, int32_t wc,
                    int32_t nc, int32_t conv_out_form, float* partial, int32_t parts, const float* inv_scale, float* dw,
                    float* narrow_sum, void* stream);

/* Training forward of the attention core on the tcgen05 kernel: like dsg_attention, and also writes
 * lse float[n][heads][tokens], the log2-sum-exp of each query's scaled scores, for dsg_attention_bwd.
 * Shapes: dsg_attention_train_tc_ok(tokens, head_dim) == 1 (head_dim 8, tokens a multiple of 128 in [128, 2048]). */
int dsg_attention_train_tc_ok(int32_t tokens, int32_t head_dim);
int dsg_attention_train(const void* qkv, void* out, float* lse, int32_t n, int32_t tokens, int32_t heads,
                        int32_t head_dim, void* stream);
/* Backward of the attention core (head_dim 8 only): dqkv h16 [n][tokens][3*c] from qkv, the forward output `out` and
 * its gradient `dout` (both h16 [n][tokens][c]).
 *   lse != NULL (from dsg_attention_train): tcgen05 kernel (two passes of recomputed probabilities, TMEM operands);
 *   lse == NULL: CUDA-core flash-style kernels, any token count; ws: float[2 * n * heads * tokens] scratch
 *   (ws may be NULL when lse is given and the shape is inside the tcgen05 kernel). */
int dsg_attention_bwd(const void* qkv, const void* out, const void* dout, void* dqkv, float* ws, const float* lse,
                      int32_t n, int32_t tokens, int32_t heads, int32_t head_dim, void* stream);

/* All finalisers of one backward pass in one launch.  A job is one dsg_gn_bwd_params (comps = 2: src = partial +
 * chunks * c * 2, parts = 1, sample_stride = (chunks + 1) * c * 2, part_stride = 0, out0 = d beta, out1 = d gamma) or one
 * dsg_colsum_finalize (comps = 1: src = partial, sample_stride = parts * c, part_stride = c, out0 / out0b = totals,
 * per_n = optional per-sample sums); block_begin = first block of the job in the flat grid, ceil(c / 32) blocks each,
 * jobs sorted by block_begin.  Same arithmetic and order as the single calls (bit-identical results). */
typedef struct dsg_reduce_job {
  const float* src;
  int32_t n, parts, c, comps;
  int64_t sample_stride, part_stride;
  float* per_n;
  int32_t per_n_stride, per_n_off;
  const float* inv_scale;
  float* out0;
  float* out0b;
  float* out1;
  int32_t block_begin, pad_;
} dsg_reduce_job;
int dsg_reduce_rows_batched(const dsg_reduce_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream);

/* Global gradient norm over one flat fp32 buffer: out3 = { ||g|| * inv_loss_scale, coefficient that unscales and clips
 * (inv_loss_scale * min(1, max_norm / (norm + 1e-6)); max_norm <= 0 disables clipping), 1 if the norm is inf/nan }.
 * partial: double[parts] scratch.  Deterministic (fixed-order sums). */
int dsg_grad_norm(const float* g, int64_t numel, double* partial, int32_t parts, float inv_loss_scale, float max_norm,
                  float* out3, void* stream);
/* torch.optim.AdamW step (decoupled weight decay, bias correction, no amsgrad) over flat buffers; step >= 1.
 * ctl (may be NULL): device float[3] from dsg_grad_norm — gradients are multiplied by ctl[1] and the whole update is
 * skipped when ctl[2] != 0 (GradScaler semantics). */
int dsg_adamw_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int32_t step, const float* ctl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSG_B200_H_ */
