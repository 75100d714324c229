/* C ABI of the B200-native molecular-VAE ELBO hot path (libmvae_b200.so).
 *
 * The reference (aclyde11/molecular-VAE) has no FFI layer: its hot path sits directly behind
 * torch.nn.Module.  These entry points are what a binding for that path attaches to;
 * each names the reference code it replaces.  Plain pointers and sizes only, no torch types.  All
 * pointers are DEVICE pointers unless stated; the caller owns every buffer including the workspace;
 * nothing here allocates device memory or synchronises (except mvae_cfgb_read_error); all work is
 * enqueued on the given stream of the current device.  Return value: 0 or a negative MVAE_ERR_* code
 * (mvae_strerror).  There is no CPU fallback: without a B200 the calls fail with MVAE_ERR_CUDA.
 */
#ifndef MVAE_B200_H_
#define MVAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mvae_stream_t; /* == cudaStream_t */

#define MVAE_OK 0
#define MVAE_ERR_INVALID (-1)
#define MVAE_ERR_WORKSPACE (-2)
#define MVAE_ERR_CUDA (-3)
#define MVAE_ERR_UNSUPPORTED (-4)
#define MVAE_ERR_DRIVER (-5)

#define MVAE_PREC_FP32 0 /* check mode: fp32 storage and CUDA-core fp32 FMA everywhere            */
#define MVAE_PREC_BF16 1 /* bf16 activations/weights into tcgen05 tensor cores, fp32 accumulation  */

const char* mvae_strerror(int rc);
const char* mvae_last_cuda_error(void);
/* number of kernels / memsets this library has enqueued since the last reset (bench.py gpu_launches) */
long long mvae_launch_count(void);
void mvae_reset_launch_count(void);
/* Per-kernel timing for bench.py's roofline: while enabled, every DIRECT (not graph-captured) launch of the persistent
 * recurrence kernels is bracketed by CUDA events on its stream.  tag 0 = forward sweep, 1 = BPTT sweep (one launch = all
 * T steps of one GRU layer).  mvae_profile_read synchronises on the recorded events.                                  */
int mvae_profile_enable(int on);
int mvae_profile_read(int tag, float* total_ms, int* launches);

/* ---- "Config B": the canonical conv / latent-Z / L x H GRU model -------------------------------
 * models2d.py:8-52 (class VAE: encode / reparametrize / decode / forward) with the latent widened
 * to `latent`, trained with loss_function of train.py:31-38.                                       */
typedef struct mvae_cfgb_desc {
  int32_t batch;     /* B molecules in this call                                                   */
  int32_t seq_len;   /* T = 120: conv1 in-channels and GRU steps (models2d.py:12,42)                */
  int32_t charset;   /* C = 35 (<= 64)                                                             */
  int32_t latent;    /* Z = 292                                                                    */
  int32_t hidden;    /* H = 501                                                                    */
  int32_t layers;    /* L = 3 (1..4)                                                               */
  int32_t fc0;       /* 435 (models2d.py:15)                                                       */
  int32_t precision; /* MVAE_PREC_*                                                                */
  int32_t train;     /* 1: z = mu + eps_scale*eps*exp(logvar/2); 0: z = mu (models2d.py:32-38)      */
  float max_len;     /* multiplier of the BCE mean (script global `max_len`, train.py:35)          */
  float eps_scale;   /* 1.0 for models2d.py:34-35; 1e-2 for models.py Lambda (models.py:85,92)      */
} mvae_cfgb_desc;

/* Parameter / gradient tensors are passed as HOST arrays of fp32 DEVICE pointers in state_dict order
 * (SURVEY.md A.1), unpadded, row-major, exactly the reference's shapes:
 *   0 conv1d1.weight (9,T,9)  1 conv1d1.bias  2 conv1d2.weight (9,9,9)  3 conv1d2.bias
 *   4 conv1d3.weight (10,9,11) 5 conv1d3.bias 6 fc0.weight (fc0,10*(C-26)) 7 fc0.bias
 *   8 fc11.weight (Z,fc0) 9 fc11.bias 10 fc12.weight 11 fc12.bias 12 fc2.weight (Z,Z) 13 fc2.bias
 *   14+4l gru.weight_ih_l (3H,in)  15+4l gru.weight_hh_l (3H,H)  16+4l gru.bias_ih_l  17+4l gru.bias_hh_l
 *   14+4L fc3.weight (C,H)  15+4L fc3.bias                                                         */
#define MVAE_CFGB_NUM_PARAMS(layers) (16 + 4 * (layers))

size_t mvae_cfgb_workspace_bytes(const mvae_cfgb_desc* d);

/* Fused ELBO step: forward + loss + backward of one batch (replaces train.py:98-101 =
 * model(data); loss_function(...); loss.backward()).  ids: u8 (B,T) character ids (the argmax of the
 * reference's one-hot input, data_loader.py:26-31); eps: fp32 (B,Z) standard-normal draws (ignored when
 * train == 0).  grads are OVERWRITTEN.  out_scalars (device, 4 floats): loss, max_len*BCE, KL (swapped
 * form of train.py:36-37), number of molecules whose argmax reconstruction is exact (train.py:110-112).
 * mu_out / logvar_out: optional fp32 (B,Z).                                                        */
int mvae_cfgb_elbo_step(const mvae_cfgb_desc* d, const float* const* params, float* const* grads,
                        const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out,
                        float* logvar_out, void* workspace, size_t workspace_bytes, mvae_stream_t stream);

/* Same work captured once into a CUDA graph (all pointers must stay valid and fixed); launch replays it. */
typedef struct mvae_graph mvae_graph;
int mvae_cfgb_elbo_step_graph_create(const mvae_cfgb_desc* d, const float* const* params, float* const* grads,
                                     const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out,
                                     float* logvar_out, void* workspace, size_t workspace_bytes,
                                     mvae_graph** out_graph);
/* The same step cut into `layers` phases for data-parallel runs (train_distributed.py:72): after phase p
 * the gradients of one contiguous bucket of the state_dict order are final and can be all-reduced while the next phase
 * runs.  phase 0: forward, loss, head (fc3) and the top GRU layer; phase k: GRU layer L-1-k; phase L-1 additionally the
 * latent layers and the encoder, and writes out_scalars.  Running phases 0..L-1 in order equals mvae_cfgb_elbo_step.   */
int mvae_cfgb_elbo_step_phase(const mvae_cfgb_desc* d, const float* const* params, float* const* grads,
                              const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out,
                              float* logvar_out, void* workspace, size_t workspace_bytes, int phase,
                              mvae_stream_t stream);
int mvae_cfgb_elbo_step_phase_graph_create(const mvae_cfgb_desc* d, const float* const* params, float* const* grads,
                                           const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out,
                                           float* logvar_out, void* workspace, size_t workspace_bytes, int phase,
                                           mvae_graph** out_graph);
int mvae_graph_launch(mvae_graph* g, mvae_stream_t stream);
long long mvae_graph_num_kernel_nodes(const mvae_graph* g);
void mvae_graph_destroy(mvae_graph* g);

/* Drop-in forward of models2d.VAE.forward (models2d.py:49-52): probs fp32 (B,T,C) = softmax over the
 * charset, mu, logvar fp32 (B,Z).  Leaves the activations needed by mvae_cfgb_backward in `workspace`. */
int mvae_cfgb_forward(const mvae_cfgb_desc* d, const float* const* params, const uint8_t* ids, const float* eps,
                      float* probs, float* mu, float* logvar, void* workspace, size_t workspace_bytes,
                      mvae_stream_t stream);
/* Backward of that forward for arbitrary upstream gradients (autograd of the reference module):
 * dprobs (B,T,C), dmu, dlogvar (B,Z) fp32, any of them may be NULL (= zero).  grads are OVERWRITTEN.   */
int mvae_cfgb_backward(const mvae_cfgb_desc* d, const float* const* params, float* const* grads, const uint8_t* ids,
                       const float* eps, const float* dprobs, const float* dmu, const float* dlogvar, void* workspace,
                       size_t workspace_bytes, mvae_stream_t stream);

/* Greedy decode of fixed latents z fp32 (B,Z) (train.py:110 / train_sample.py:31-33 applied to the Config-B
 * decoder, models2d.py:40-47): ids_out u8 (B,T) = argmax over the charset (ties -> lowest id);
 * probs_out optional fp32 (B,T,C).                                                                  */
int mvae_cfgb_decode_greedy(const mvae_cfgb_desc* d, const float* const* params, const float* z, uint8_t* ids_out,
                            float* probs_out, void* workspace, size_t workspace_bytes, mvae_stream_t stream);

/* featurizer.py:8-24 / data_loader.py:26-31 one-hot (rows, C) fp32 -> u8 ids; *not_onehot (device int) is
 * set to 1 when a row is not exactly one-hot.                                                        */
int mvae_onehot_to_ids(const float* onehot, long long rows, int charset, uint8_t* ids, int* not_onehot,
                       mvae_stream_t stream);

/* Device-side error flag of the last call that used `workspace` (tcgen05 pipeline watchdog).  Synchronises
 * the stream.  Writes 0/1 to *flag (host).                                                            */
int mvae_cfgb_read_error(const mvae_cfgb_desc* d, void* workspace, size_t workspace_bytes, int* flag,
                         mvae_stream_t stream);

/* ---- "Config A": models.py exactly as shipped ----------------------------------------------------------------
 * models.py:97-165 (MolecularVAE = MolEncoder :109-135 [Embedding -> LSTM -> 3 x ConvSELU(k18) -> Linear+SELU -> Lambda
 * :80-94] + MolDecoder :148-165 [Linear+SELU -> Repeat -> LSTM -> Linear -> Softmax]), trained with loss_function of
 * train.py:31-38.  The conv widths (120, 64, 64), kernel size 18 and dense_1 width 512 are the reference's literals.     */
typedef struct mvae_cfga_desc {
  int32_t batch;      /* B                                                                                            */
  int32_t seq_len;    /* T = 120 = `i`: LSTM steps AND conv_1 in-channels (models.py:118)                              */
  int32_t charset;    /* C = 35 (<= 64)                                                                               */
  int32_t embed;      /* word_embedding_size = 30 (models.py:111)                                                     */
  int32_t enc_hidden; /* h_size = 72  (>= 52 so that three k=18 convolutions fit)                                     */
  int32_t enc_layers; /* num_lstm = 3 (1..4)                                                                          */
  int32_t latent;     /* o = 292                                                                                      */
  int32_t dec_hidden; /* 1024 (models.py:150), multiple of 64                                                         */
  int32_t dec_layers; /* num_gru = 4 (1..4)                                                                           */
  int32_t precision;  /* MVAE_PREC_*                                                                                  */
  float max_len;      /* multiplier of the BCE mean (script global, 128 in train.py:43)                               */
  float eps_scale;    /* Lambda.scale = 1e-2 (models.py:82,92)                                                        */
} mvae_cfga_desc;
/* parameters / gradients: host arrays of fp32 device pointers in state_dict order:
 *   0 encoder.embedding.weight (C,E) | 1+4l.. encoder.gru.{weight_ih_l, weight_hh_l, bias_ih_l, bias_hh_l} (4*EH rows, i,f,g,o)
 *   then encoder.conv_{1,2,3}.0.{weight,bias}, encoder.dense_1.0.{weight,bias}, encoder.lmbd.z_mean.{weight,bias},
 *   encoder.lmbd.z_log_var.{weight,bias}, decoder.latent_input.0.{weight,bias}, decoder.gru.{...}_l (4 per layer),
 *   decoder.decoded_mean.module.0.{weight,bias}                                                                       */
#define MVAE_CFGA_NUM_PARAMS(enc_layers, dec_layers) (17 + 4 * (enc_layers) + 4 * (dec_layers))
size_t mvae_cfga_workspace_bytes(const mvae_cfga_desc* d);
/* Fused step = train.py:98-101 (model(data); loss_function(...); loss.backward()) for the shipped model.  ids u8 (B,T);
 * eps fp32 (B,Z) standard-normal draws (the reference draws them on the CPU generator, models.py:92).  grads are
 * OVERWRITTEN.  out_scalars as for mvae_cfgb_elbo_step.                                                               */
int mvae_cfga_elbo_step(const mvae_cfga_desc* d, const float* const* params, float* const* grads, const uint8_t* ids,
                        const float* eps, float* out_scalars, float* mu_out, float* logvar_out, void* workspace,
                        size_t workspace_bytes, mvae_stream_t stream);
int mvae_cfga_elbo_step_graph_create(const mvae_cfga_desc* d, const float* const* params, float* const* grads,
                                     const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out,
                                     float* logvar_out, void* workspace, size_t workspace_bytes, mvae_graph** out_graph);
/* MolecularVAE.forward (models.py:104-106): probs fp32 (B,T,C), mu, logvar fp32 (B,Z); and its autograd backward.        */
int mvae_cfga_forward(const mvae_cfga_desc* d, const float* const* params, const uint8_t* ids, const float* eps,
                      float* probs, float* mu, float* logvar, void* workspace, size_t workspace_bytes,
                      mvae_stream_t stream);
int mvae_cfga_backward(const mvae_cfga_desc* d, const float* const* params, float* const* grads, const uint8_t* ids,
                       const float* eps, const float* dprobs, const float* dmu, const float* dlogvar, void* workspace,
                       size_t workspace_bytes, mvae_stream_t stream);
/* MolDecoder.forward on given latents (train_sample.py:31-33: model.decoder(z) -> argmax): ids_out u8 (B,T), probs_out
 * optional fp32 (B,T,C).                                                                                              */
int mvae_cfga_decode(const mvae_cfga_desc* d, const float* const* params, const float* z, uint8_t* ids_out,
                     float* probs_out, void* workspace, size_t workspace_bytes, mvae_stream_t stream);
int mvae_cfga_read_error(const mvae_cfga_desc* d, void* workspace, size_t workspace_bytes, int* flag,
                         mvae_stream_t stream);

/* ---- MOSES-style character VAE ------------------------------------------------------------------------------
 * mosesvae.py:27-199 (class VAE: forward :126-140, forward_encoder :142-164, forward_decoder :166-199).       */
typedef struct mvae_moses_desc {
  int32_t batch;      /* B sequences, sorted by length descending as the reference's collate does             */
  int32_t max_len;    /* T: padded length of `ids` in this call (= longest sequence incl. bos/eos)             */
  int32_t vocab;      /* V = len(vocab) <= 256 (chars / SELFIES symbols + bos,eos,pad,unk; vocab.py:24; ids are u8); embedding V x V; for V > 64 the token-table projections are materialised instead of looked up inside the sweeps */
  int32_t d_z;        /* 160 (mosesvae.py:39)                                                                 */
  int32_t q_hidden;   /* 256 encoder GRU (mosesvae.py:33), multiple of 64                                     */
  int32_t d_hidden;   /* 512 decoder GRU (mosesvae.py:40), multiple of 64                                     */
  int32_t d_layers;   /* 3                                                                                    */
  int32_t mlp_hidden; /* 256: hidden width of the q_mu / q_logvar MLPs (mosesvae.py:68-69)                     */
  int32_t pad_id;     /* vocab.pad: ignore_index of the CE and padding_idx of the embedding                   */
  int32_t precision;  /* MVAE_PREC_*                                                                          */
  float kl_weight;    /* the scalar differentiated is kl_weight*kl + recon_weight*recon                        */
  float recon_weight; /* (moses_train_distrib_logp.py:302-306 uses kl_weight*kl + recon, :335 recon only)      */
  /* mosesfile.py variant (mosesfile.py:6-157; BASELINE config 4): */
  int32_t q_bidir;        /* 1: bidirectional encoder GRU (mosesfile.py:21-28), heads read cat(h_fwd, h_bwd) (:115-116) */
  int32_t q_linear_heads; /* 1: q_mu / q_logvar are single Linear(Hq*(1+bidir), d_z) (mosesfile.py:31-32); 0: the 2-layer MLPs */
  /* train-mode dropout between decoder GRU layers (nn.GRU(dropout=0.2), mosesvae.py:38,78): element (l, t, b, j) of the
   * output of layer l < L-1 is kept iff u16(dropout_seed, l, t, b, j) >= d_dropout and scaled by 1/(1-d_dropout); the
   * counter-based 16-bit uniform (one 64-bit hash per four consecutive units) is restated in
   * oracle/moses_oracle.dropout_masks so a parity run can inject the same mask.  0 = off. */
  float d_dropout;
  uint32_t dropout_seed;
} mvae_moses_desc;
/* parameters / gradients: host arrays of fp32 device pointers in this order (reference shapes, row-major):
 *   0 x_emb.weight (V,V) | 1-4 encoder_rnn.{weight_ih_l0 (3Hq,V), weight_hh_l0, bias_ih_l0, bias_hh_l0}
 *   5-8 q_mu.{0.weight (mlp,Hq), 0.bias, 2.weight (d_z,mlp), 2.bias} | 9-12 q_logvar.{...}
 *   13+4l.. decoder_rnn.{weight_ih_l (3Hd, V+d_z | Hd), weight_hh_l, bias_ih_l, bias_hh_l}
 *   13+4L decoder_lat.weight (Hd,d_z), +1 decoder_lat.bias, +2 decoder_fc.weight (V,Hd), +3 decoder_fc.bias
 * With q_bidir the four encoder_rnn.*_l0_reverse tensors follow the forward ones; with q_linear_heads each head is
 * {weight (d_z, Hq*(1+bidir)), bias} (state_dict order of mosesfile.VAE).                                        */
#define MVAE_MOSES_NUM_PARAMS(layers) (17 + 4 * (layers))
#define MVAE_MOSESFILE_NUM_PARAMS(layers) (17 + 4 * (layers))   /* 1 + 8 + 4 + 4L + 4 */
size_t mvae_moses_workspace_bytes(const mvae_moses_desc* d);
/* One step of VAE.forward (+ backward when grads != NULL; dropout is the identity).  ids: u8 (B,T) right-padded
 * with pad_id; lengths: int32 (B) (incl. bos/eos); eps: fp32 (B,d_z).  out_scalars (device, 4 floats):
 * kl_weight*kl + recon_weight*recon, kl, recon, number of non-pad targets.  z_out, logvar_out (B,d_z) and
 * y_out (B,T,V; the logits the reference returns, decoder_fc bias at padded positions) are optional.
 * lengths_host (optional HOST copy of `lengths`): enables torch's packed-sequence batch sizes -- step t of every
 * recurrence only processes the sequences that are still running (a prefix of the length-sorted batch).            */
int mvae_moses_step(const mvae_moses_desc* d, const float* const* params, float* const* grads, const uint8_t* ids,
                    const int32_t* lengths, const int32_t* lengths_host, const float* eps, float* out_scalars,
                    float* z_out, float* logvar_out, float* y_out, void* workspace, size_t workspace_bytes,
                    mvae_stream_t stream);
/* mvae_moses_step with two more inputs (grads must be given when phase >= 0):
 *   dz_ext  optional fp32 (B,d_z): gradient wrt z from a consumer of z outside the VAE (autograd of a property head,
 *           trainbinding.py:216-217); added to the decoder's gradient wrt z before the reparametrisation backward.
 *   phase   -1 = the whole step.  0..d_layers-1 = the part of the step whose gradients become final in phase p, in
 *           backward order (the data-parallel bucket order, as mvae_cfgb_elbo_step_phase): phase 0 = forward + loss +
 *           decoder_fc + top decoder layer; phase k = decoder layer d_layers-1-k; the last phase additionally the layer-0
 *           input weights, decoder_lat, the mu / logvar heads, the encoder GRU(s) and x_emb.  Run all phases in order on
 *           one stream with the same buffers; between two phases the finished slice can be all-reduced
 *           (moses_train_distrib.py:30-42,176,191 -- the reference's commented-out DDP).                               */
int mvae_moses_step_ex(const mvae_moses_desc* d, const float* const* params, float* const* grads, const uint8_t* ids,
                       const int32_t* lengths, const int32_t* lengths_host, const float* eps, const float* dz_ext,
                       float* out_scalars, float* z_out, float* logvar_out, float* y_out, void* workspace,
                       size_t workspace_bytes, int phase, mvae_stream_t stream);
/* VAE.sample (mosesvae.py:214-262; hugesample.py:28): autoregressive decode of B latents z fp32 (B,d_z) for
 * max_len-1 steps (desc->max_len = the sampler's max_len, 100 in the reference).  mode 0 = greedy argmax (ties ->
 * lowest id; the bit-exact parity mode), mode 1 = multinomial over softmax(y/temp) with a counter-based generator
 * (seed, sequence, step).  ids_out u8 (B,max_len): bos, generated ids up to and including eos, pad after;
 * lengths_out int32 (B) = the reference's end_pads (eos index + 1, or max_len).  All rows run every step.
 * seed_device (optional, DEVICE pointer): when given the generator seed is read from it at run time instead of `seed`,
 * so a CUDA graph captured around this call can be replayed with fresh draws.                                        */
int mvae_moses_sample(const mvae_moses_desc* d, const float* const* params, const float* z, int bos_id, int eos_id,
                      int mode, float temp, unsigned long long seed, const unsigned long long* seed_device,
                      uint8_t* ids_out, int32_t* lengths_out, void* workspace, size_t workspace_bytes,
                      mvae_stream_t stream);
/* The sampler captured once into a CUDA graph over FIXED buffers (params, z, seed_device, ids_out, lengths_out,
 * workspace); replay with mvae_graph_launch after writing fresh z / seed into the same buffers.  The weights are
 * re-read from `params` at every replay.  Replaces the 99-iteration Python loop of mosesvae.py:239-251.          */
int mvae_moses_sample_graph_create(const mvae_moses_desc* d, const float* const* params, const float* z, int bos_id,
                                   int eos_id, int mode, float temp, const unsigned long long* seed_device,
                                   uint8_t* ids_out, int32_t* lengths_out, void* workspace, size_t workspace_bytes,
                                   mvae_graph** out_graph);
int mvae_moses_read_error(const mvae_moses_desc* d, void* workspace, size_t workspace_bytes, int* flag,
                          mvae_stream_t stream);

/* ---- BindingModel property head (mosesvae.py:6-25; callers moses_train_distrib.py:274, trainbinding.py:216) ----------
 * Linear(Z,256) -> BatchNorm1d(256) -> Tanh -> Linear(256,256) -> ReLU -> Linear(256,64) -> BatchNorm1d(64) -> ReLU -> Linear(64,1)
 * on latents z fp32 (B,Z).  params / grads: host arrays of 12 fp32 device pointers in state_dict order
 *   binding_model.{0.weight (256,Z), 0.bias, 1.weight, 1.bias, 3.weight (256,256), 3.bias, 5.weight (64,256), 5.bias,
 *                  6.weight, 6.bias, 8.weight (1,64), 8.bias};
 * running: 4 device pointers {1.running_mean, 1.running_var, 6.running_mean, 6.running_var}, updated in train mode
 * (momentum bn_momentum, unbiased variance) and read in eval mode.  forward leaves the activations backward needs in
 * `workspace`; out fp32 (B) = the module's (B,1) output; dout fp32 (B); dz optional fp32 (B,Z).  grads are OVERWRITTEN.  */
typedef struct mvae_binding_desc {
  int32_t batch;
  int32_t z_size;    /* 128 (mosesvae.py:7) */
  int32_t train;     /* BatchNorm1d mode */
  float bn_eps;      /* 1e-5 */
  float bn_momentum; /* 0.1 */
} mvae_binding_desc;
size_t mvae_binding_workspace_bytes(const mvae_binding_desc* d);
/* The VAE step with the property head on z in ONE call: the historical `VAE.forward(x, binding) -> (kl, recon,
 * binding_loss, z)` that moses_train_distrib.py:274, trainbinding.py:216 and mosesanalyize.py:192 call (no shipped class
 * implements it; BASELINE.json configs[3] "with the logP property head").  binding_loss = binding_weight *
 * mean_b (BindingModel(z)_b - target_b)^2 with target fp32 (B) (the min-max scaled score / logP of
 * moses_train_distrib.py:143); the differentiated scalar is kl_weight*kl + recon_weight*recon + binding_loss, so the
 * head's gradient wrt z reaches the encoder.  bparams / bgrads / brunning as for mvae_binding_forward / _backward
 * (bgrads overwritten); out_scalars as mvae_moses_step; binding_loss_out: 1 device float; `extra`: device scratch of
 * mvae_moses_joint_extra_bytes(d) bytes (256-byte aligned); phase as mvae_moses_step_ex (the head runs in phase 0).    */
size_t mvae_moses_joint_extra_bytes(const mvae_moses_desc* d);
int mvae_moses_joint_step(const mvae_moses_desc* d, const float* const* params, float* const* grads, const uint8_t* ids,
                          const int32_t* lengths, const int32_t* lengths_host, const float* eps,
                          const mvae_binding_desc* bd, const float* const* bparams, float* const* bgrads,
                          float* const* brunning, const float* target, float binding_weight, float* out_scalars,
                          float* binding_loss_out, float* z_out, void* workspace, size_t workspace_bytes,
                          void* binding_workspace, size_t binding_workspace_bytes, void* extra, size_t extra_bytes,
                          int phase, mvae_stream_t stream);
int mvae_binding_forward(const mvae_binding_desc* d, const float* const* params, float* const* running, const float* z,
                         float* out, void* workspace, size_t workspace_bytes, mvae_stream_t stream);
int mvae_binding_backward(const mvae_binding_desc* d, const float* const* params, float* const* grads, const float* z,
                          const float* dout, float* dz, void* workspace, size_t workspace_bytes, mvae_stream_t stream);

/* ---- text assembly of decoded / sampled ids (mosesvae.py:258-262 -> vocab.py:62-73 ids2string; hugesample.py:31-35;
 * featurizer.py:26-37) -----------------------------------------------------------------------------------------------
 * ids u8 (B,L); lengths int32 (B) or NULL (every row has L tokens).  table: 256 token texts of tok_stride bytes each,
 * tok_len u8[256] their lengths (0 = emit nothing).  rem_first_id / rem_last_id: drop that id when it is the first /
 * last token of a row (ids2string's rem_bos / rem_eos), -1 = keep.  strip: drop leading / trailing single-space tokens
 * (featurizer.py:37).  Output: row b is out_bytes[out_offsets[b] .. out_offsets[b+1]) (bytes beyond `capacity` are
 * dropped, the offsets stay exact); scratch_row_len: int32 (B).  One D2H of offsets + bytes replaces B Python loops.     */
int mvae_ids_to_text(const uint8_t* ids, const int32_t* lengths, int B, int L, const uint8_t* table, int tok_stride,
                     const uint8_t* tok_len, int rem_first_id, int rem_last_id, int strip, uint8_t* out_bytes,
                     long long capacity, int32_t* out_offsets, int32_t* scratch_row_len, mvae_stream_t stream);

/* ---- input featurisation on the device (featurizer.py:8-24 OneHotFeaturizer.featurize,
 * data_loader.py:26-31): `text` holds the B strings back to back (bytes), row b = text[offsets[b] .. offsets[b+1]);
 * lut u8[256] maps a byte to its charset index (255 = not in the charset).  ids_out u8 (B,T): the string's ids, then pad_id
 * (the charset index of ' ', i.e. ljust(T), featurizer.py:19-20).  bad_flag (device int32): bit 0 = a character outside
 * the charset (featurizer.py:17 raises ValueError there), bit 1 = a string longer than T.  The model entry points take
 * these ids directly, so the (B,T,C) one-hot never exists on the host or crosses PCIe.                                   */
int mvae_text_to_ids(const uint8_t* text, const int32_t* offsets, int B, int T, const uint8_t* lut, int pad_id,
                     uint8_t* ids_out, int32_t* bad_flag, mvae_stream_t stream);

/* ---- optimiser step on flat fp32 buffers (train.py:102-104, train_distributed.py:91-94) ------------------
 * Global-norm clipping = torch.nn.utils.clip_grad_norm(params, max_norm) over ONE flat gradient buffer (the layout
 * molecular-vae_b200/ddp.py uses); the clip coefficient min(1, max_norm/(norm+1e-6)) stays on the device at
 * (float*)((char*)scratch16 + 8) so no host sync is needed; apply_scale=0 leaves the gradients untouched and lets the
 * optimiser kernels fold the coefficient in.  Adam = torch.optim.Adam (train.py:81), SGD = torch.optim.SGD with
 * momentum (train_distributed.py:73).                                                                      */
int mvae_clip_grad_norm(float* grads, long long n, float max_norm, void* scratch16, float* norm_out, int apply_scale,
                        mvae_stream_t stream);
int mvae_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, const float* clip_coef,
                   mvae_stream_t stream);
int mvae_sgd_momentum_step(float* params, const float* grads, float* momentum_buf, long long n, float lr,
                           float momentum, float weight_decay, int first_step, const float* clip_coef,
                           mvae_stream_t stream);

/* ---- building blocks, exported for the parity tests ---------------------------------------------- */
/* D[M,N] (+)= A*B (+bias[N]); bf16 operands on the tcgen05 path.  a_mn_major: A stored [K][M];
 * b_mn_major: B stored [K][N] (else [N][K]).  out fp32 or bf16.                                       */
int mvae_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                   void* D, long long ldd, int d_is_bf16, int accumulate, const float* bias, int M, int N, int K,
                   int tile_n, int splits, int* err_flag, mvae_stream_t stream);
/* fp32 CUDA-core GEMM with explicit strides: A(m,k)=A[m*sam+k*sak], B(k,n)=B[k*sbk+n*sbn].            */
int mvae_sgemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
               long long ldc, int M, int N, int K, const float* bias, int act, int accumulate, int splits,
               mvae_stream_t stream);
/* The same product on the tensor cores at fp32-class accuracy (bf16 mode's path for the small latent / encoder Linears:
 * models2d.py:28-29,41, mosesvae.py:68-69,95): each operand is split into bf16 hi + lo parts and A*B ~ Ah*Bh + Ah*Bl + Al*Bh
 * runs as ONE tcgen05 GEMM over a 3x longer contraction (fp32 accumulation, ~2^-17 relative per product).  `scratch`
 * (mvae_sgemm_tc_scratch_bytes(max rows, max cols of either operand), 256-byte aligned) holds the converted operands.
 * Strides must be contiguous along k or along m / n; returns MVAE_ERR_UNSUPPORTED (nothing enqueued) otherwise, for tiny
 * products (M*N*K < 2^22) or when the scratch is too small.  act: 0 none, 1 SELU, 2 ReLU (after the bias).                 */
int mvae_sgemm_tc(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
                  long long ldc, int M, int N, int K, const float* bias, int act, int accumulate, void* scratch,
                  size_t scratch_bytes, int* err_flag, mvae_stream_t stream);
size_t mvae_sgemm_tc_scratch_bytes(long long rows_max, long long cols_max);
/* L2 <-> SM fabric probe (the roofline denominator bench.py reports for the persistent recurrence sweeps): `ctas` blocks of
 * 512 threads stream `buf` (`bytes`, keep it well inside the L2: 32 MiB) `passes` times with 16-byte L1-bypassing accesses.
 * mode 0: read only (bytes * passes move L2 -> SM); mode 1: the first half is copied onto the second (bytes/2 * passes read
 * plus the same written).  The caller times the launch with CUDA events.                                                  */
int mvae_l2_probe(void* buf, size_t bytes, int passes, int mode, int ctas, mvae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVAE_B200_H_ */
