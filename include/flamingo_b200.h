/* flamingo_b200.h — C ABI of libflamingo_b200.so (B200 / sm_100a kernels for the flamingo-mini hot path).
 *
 * The reference (dhansmair/flamingo-mini) has no FFI: its "operator API" for this path is the nn.Module contract of
 *   flamingo_mini/perceiver_resampler.py:99-188   PerceiverResampler(dim, depth, ...).forward(x_f)
 *   flamingo_mini/gated_cross_attention.py:135-184 GatedCrossAttentionBlock(...).forward(y, visual_features, media_locations, ...)
 *   flamingo_mini/utils.py:31-50                  FeedForward
 * The entry points below are what a ctypes/cffi binding under those modules binds; the Python
 * package flamingo_mini_b200 is exactly such a binding and keeps the reference's module names, constructor
 * keywords, forward signatures and parameter names.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless noted; plain pointers and sizes only, no torch types;
 *   - the library never allocates, frees or retains device memory: outputs, saved activations and scratch are
 *     caller-allocated (sizes from the *_bytes() helpers); all work is enqueued on the cudaStream_t passed in;
 *   - every function returns 0 on success, a negative FM_E* code otherwise; fm_last_error() gives the text
 *     (thread-local). Nothing throws, nothing falls back to another implementation;
 *   - activations are bf16 row-major unless a *_f32 flag says fp32; parameters are given twice: the fp32 master
 *     ("w_f32", used for LayerNorm affine + gates) and a bf16 shadow with the SAME element layout ("w_bf16", used
 *     as tensor-core operands); gradients are written (not accumulated) into an fp32 buffer with that layout.
 */
#ifndef FLAMINGO_B200_H_
#define FLAMINGO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* fm_stream_t; /* == cudaStream_t */

enum {
  FM_OK = 0,
  FM_EINVAL = -1,      /* bad shape / unsupported configuration */
  FM_ECUDA = -2,       /* CUDA runtime / driver error */
  FM_EUNSUPPORTED = -3 /* device is not sm_100 */
};

enum { FM_ACT_GELU = 0, FM_ACT_SQRELU = 1, FM_ACT_RELU = 2 }; /* utils.py:36-40 */

int fm_version(void);                /* ABI version, currently 1 */
const char* fm_last_error(void);     /* host pointer, thread-local */
unsigned int fm_device_error(void);  /* device-side watchdog word (0 = none); synchronises the device */

/* options: FM_OPT_SIDE_STREAM (default 1) - weight-gradient GEMMs are issued on a library-owned side stream forked from /
 * joined into the caller's stream (parallel branches under graph capture); 0 keeps every kernel on the caller's stream.
 * Keys >= 1 are scheduling switches; they never change a result beyond floating-point summation order (defaults in
 * parentheses; the measured effect of each at C2 is in profiles/r02_validate_next/summary.log):
 *   FM_OPT_GEMM_GROUP (1)      independent GEMMs of one phase (dWout+dWq+dWkv, q+kv, dyn+dvis) share one persistent launch
 *   FM_OPT_EPI_PREFETCH (0)    (historic name) TMA L2 prefetch of the OPERAND boxes two ring depths ahead of the producer for contractions
 *                              of >= 16 k-blocks.  Off: measured slower (the prefetch operations compete with the producer's own loads);
 *                              epilogue inputs are not prefetched into L2 any more either: they arrive by per-warp TMA loads
 *   FM_OPT_ALPHA_FROM_DW2 (1)  d(alpha_ffw) = sum(W2 * dW2_ungated) from the dW2 epilogue instead of sum(dH * h) in DACT
 *   FM_OPT_PDL (1)             programmatic dependent launch: a kernel's prologue overlaps its predecessor's tail
 *   FM_OPT_LN_REDUCE_SIDE (1)  the dgamma/dbeta fold of LayerNorm backward runs on the side stream
 *   FM_OPT_DATTN_FROM_GEMM (1) d(alpha_attn) = sum(dO_ungated * O) from the epilogue of the dO GEMM (fp32 accumulators) instead of a
 *                              separate dot-product kernel over the bf16-rounded dO
 *   FM_OPT_ATTN_TMEM_COMPACT (1) backward attention cores allocate 256 instead of 512 TMEM columns (accumulators that are never live
 *                              together share columns), so the two CTAs an SM holds run side by side instead of one after the other
 *   FM_OPT_DEFER_JOIN (0)      fm_xattn_bwd returns WITHOUT joining the side stream: its weight-gradient GEMMs (and LayerNorm folds)
 *                              keep running while the caller's stream goes on with whatever comes next (the frozen LM block's
 *                              backward).  Contract for the caller while this is on: every buffer passed to that call (saved,
 *                              scratch, dy_out, visual features, the gradient arena) stays allocated, and nothing reads the
 *                              parameter gradients, until fm_side_join(stream) has been called on the same stream (once per step
 *                              is enough; it is also what lets a CUDA-graph capture end)
 *   FM_OPT_DW_SPLITK (1)       weight-gradient GEMMs with few output tiles use 256-wide tiles and a PARALLEL split over K (every range
 *                              adds its partial tile with a TMA reduce-add into the pre-zeroed gradient): fewer operand bytes per
 *                              FLOP through the L2 and a unit for every SM; summation order across ranges is not fixed
 *   FM_OPT_SM_RESERVE (0)      number of SMs the persistent GEMM grids leave free (value, not a flag): under data
 *                              parallelism NCCL's CTAs occupy SMs for the length of a collective, and a persistent grid of
 *                              one CTA per SM would otherwise run its last CTAs as a second wave */
enum {
  FM_OPT_SIDE_STREAM = 0, FM_OPT_GEMM_GROUP = 1, FM_OPT_EPI_PREFETCH = 2, FM_OPT_ALPHA_FROM_DW2 = 3, FM_OPT_PDL = 4,
  FM_OPT_LN_REDUCE_SIDE = 5, FM_OPT_SM_RESERVE = 6, FM_OPT_DATTN_FROM_GEMM = 7, FM_OPT_ATTN_TMEM_COMPACT = 8,
  FM_OPT_DEFER_JOIN = 9, FM_OPT_DW_SPLITK = 10, FM_OPT_COUNT = 11
};
int fm_set_option(int key, int value);

/* launch accounting / in-situ kernel timing (bench.py): fm_launch_count() = kernels launched so far by this library;
 * after fm_profile_enable(1) every launch is bracketed by CUDA events on its stream, fm_profile_report() synchronises
 * and writes one text line per kernel tag: "tag launches total_ms flops bytes". */
unsigned long long fm_launch_count(void);
int fm_profile_enable(int on);
int fm_profile_report(char* buf /*host*/, size_t n);
/* fm_profile_enable(2): launch LOG only (no events, capturable at no cost): fm_profile_log() writes one line per launch since
 * the enable, in launch order: "tag flops bytes"; tags carry the module scope ("x/" gated xattn block, "r/" resampler) and a
 * leading '@' when the launch was recorded under stream capture.  bench.py matches this log with CUPTI kernel records of the
 * replayed CUDA graph to attribute device-side kernel durations to tags. */
int fm_profile_log(char* buf /*host*/, size_t n);

/* ------------------------------------------------------------------------------------------------ raw GEMM
 * D[m,n] = sum_k A(m,k) B(n,k); A(m,k) = a_mn ? A[k*lda+m] : A[m*lda+k], same for B. bf16 in, fp32 accumulate.
 * epi: 0 STORE  out = acc*scale*tanh(*gate) + col_bias[n]
 *      1 ACT    out = act(acc) (bf16), out2 = act'(acc) (bf16, optional)    -- Linear -> activation of FeedForward
 *      2 RESID  out = aux + tanh(*gate)*scale*acc                           -- gated / plain residual add
 *      3 DACT   out = tanh(*gate)*scale*acc*aux                              (aux = saved act'; no reduction output)
 *      0 STORE with red_out: *red_out += sum(acc * aux) with aux a bf16 [M, N] tile (d(alpha) dots)
 * Requirements: lda, ldb, ldo, ldaux, N multiples of 8; pointers 16-byte aligned.  bn = 0 lets the library pick
 * the tile width (64/128/192/256).  Gradient-shaped problems (small M x N, long K) can be split along K. */
typedef struct {
  int M, N, K;
  const void* A; long long lda; int a_mn;
  const void* B; long long ldb; int b_mn;
  int epi;
  void* out; long long ldo; int out_f32;
  void* out2; long long ldo2;
  const void* aux; long long ldaux; int aux_f32;
  const void* aux2; long long ldaux2;      /* unused since round 2 (kept so the struct layout is stable) */
  const float* col_bias;
  const float* gate;
  float* red_out;
  float scale;
  int act;
  int bn;
  int splits;          /* fp32 STORE only: > 1 = deterministic serial split-K over `splits` K ranges (needs splitk_flags); 0 = library's
                          choice; < 0 = PARALLEL split-K over |splits| ranges: the caller has ZEROED out, every range reduce-adds */
  int* splitk_flags;   /* zero-initialised ints, 16 per 128 x bn output tile (>= fm_gemm_splitk_flag_ints(M, N)); they are
                          left zero again on completion. NULL disables split-K. */
  long long* trace;    /* optional debug timeline, [min(tiles,SMs)][64] int64 (tools/gemm_trace.py); NULL in production */
} fm_gemm_desc;
size_t fm_gemm_splitk_flag_ints(int M, int N);
int fm_gemm_bf16(const fm_gemm_desc* d, fm_stream_t stream);
/* n (1..4) independent problems with the same a_mn / b_mn and epi = 0 (STORE, no split-K); results are those of n
 * fm_gemm_bf16 calls.  They run as ONE persistent launch (tiles of all problems share the SMs; FM_OPT_GEMM_GROUP = 0: n launches). */
int fm_gemm_bf16_group(const fm_gemm_desc* d, int n, fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ LayerNorm
 * nn.LayerNorm(D), eps 1e-5 (perceiver_resampler.py:52-53,187; gated_cross_attention.py:74; utils.py:46). */
int fm_layernorm_fwd(const void* x, int x_f32, const float* gamma, const float* beta, void* out, int out_f32,
                     float* mean, float* rstd, int rows, int D, fm_stream_t stream);
/* dx = LNbwd(dy) (+ dres); dgamma/dbeta written. part: scratch of fm_layernorm_bwd_scratch_bytes(D) bytes. */
size_t fm_layernorm_bwd_scratch_bytes(int D);
int fm_layernorm_bwd(const void* dy /*bf16*/, const void* x, int x_f32, const float* gamma, const float* mean,
                     const float* rstd, const void* dres, int dres_f32, void* dx, int dx_f32, float* dgamma,
                     float* dbeta, void* part, int rows, int D, fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ misc
 * text_time[b,i] = cumsum_i(media_locations[b,:]) (gated_cross_attention.py:97); int32 in/out. */
int fm_text_time(const int* media_locations, int* text_time, int B, int S, fm_stream_t stream);
int fm_cast_f32_to_bf16(const float* src, void* dst, long long n, fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ gated xattn block
 * GatedCrossAttentionBlock.forward (gated_cross_attention.py:160-184) incl. MaskedCrossAttention (:42-131) and the
 * gated FeedForward (utils.py:31-50).  dim_head must be 64 and n_visual 64; heads 1..64 (inner width I = 64*heads; the shapes
 * quoted below are those of the default 8 heads, I = 512). */
typedef struct {
  int B, S;            /* text batch, tokens */
  int D, Dv;           /* LM width, visual width */
  int n_media;         /* images per sample (keys = n_media*64) */
  int heads, dim_head; /* 1..64 (default 8), 64 */
  int ff_inner;        /* int(D*ff_mult) */
  int act;             /* FM_ACT_* */
  int y_f32;           /* dtype of y / y_out / dy_out / dy (0 = bf16) */
  int training;        /* 1: keep what backward needs */
} fm_xattn_cfg;

/* element offsets of each parameter inside the flat fp32 / bf16 / grad buffers (names: gated_cross_attention.py) */
typedef struct {
  long long attn_norm_w, attn_norm_b;  /* attn.norm.{weight,bias}      [D]            */
  long long to_q;                      /* attn.to_q.weight             [512, D]       */
  long long to_kv;                     /* attn.to_kv.weight            [1024, Dv]     */
  long long to_out;                    /* attn.to_out.weight           [D, 512]       */
  long long ffw_norm_w, ffw_norm_b;    /* ffw.0.{weight,bias}          [D]            */
  long long ffw_w1;                    /* ffw.1.weight                 [ff_inner, D]  */
  long long ffw_w2;                    /* ffw.3.weight                 [D, ff_inner]  */
  long long alpha_attn, alpha_ffw;     /* alpha_attn, alpha_ffw        [1]            */
  long long total;                     /* elements, padded to a multiple of 8         */
} fm_xattn_layout;
int fm_xattn_layout_of(const fm_xattn_cfg* cfg, fm_xattn_layout* out);
size_t fm_xattn_saved_bytes(const fm_xattn_cfg* cfg);   /* activations kept from fwd to bwd (also fwd workspace) */
size_t fm_xattn_scratch_bytes(const fm_xattn_cfg* cfg); /* backward temporaries */

/* y [B*S, D]; vis [B*n_media*64, Dv] bf16; text_time [B,S] int32; kv [B*n_media*64, 1024] bf16 is written unless
 * kv_given (cached keys/values, gated_cross_attention.py:88-92); y_out [B*S, D]. */
int fm_xattn_fwd(const fm_xattn_cfg* cfg, const float* w_f32, const void* w_bf16, const void* y, const void* vis,
                 const int* text_time, void* kv, int kv_given, void* y_out, void* saved, fm_stream_t stream);
/* dy_out: gradient w.r.t. y_out; writes dy [B*S, D], dvis [B*n_media*64, Dv] (bf16) and every parameter gradient
 * into g_f32 (fm_xattn_layout order). */
int fm_xattn_bwd(const fm_xattn_cfg* cfg, const float* w_f32, const void* w_bf16, const void* y, const void* vis,
                 const int* text_time, const void* kv, const void* saved, const void* dy_out, void* dy, void* dvis,
                 float* g_f32, void* scratch, fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ perceiver resampler
 * PerceiverResampler.forward (perceiver_resampler.py:143-188) incl. PerceiverAttentionLayer (:32-96). */
typedef struct {
  int BN;              /* batch * images */
  int T, F;            /* frames, CLIP tokens per frame (keys per layer = T*F + n_latents) */
  int Dv;
  int depth;
  int heads, dim_head; /* 1..64 (default 8), 64 */
  int n_latents;       /* 64 */
  int n_time_embeds;   /* rows of time_pos_emb (T must be <= this) */
  int ff_inner;
  int act;
  int x_f32;           /* dtype of x_f (0 = bf16) */
  int training;
} fm_resampler_cfg;

typedef struct {
  long long latents;                 /* latents            [n_latents, Dv]       */
  long long time_pos_emb;            /* time_pos_emb       [n_time_embeds,1,Dv]  */
  long long layer0;                  /* offset of layer 0; layer i at layer0 + i*layer_stride */
  long long layer_stride;
  /* offsets relative to the layer base (names: layers.{i}.0.* and layers.{i}.1.*) */
  long long norm_media_w, norm_media_b, norm_latents_w, norm_latents_b;
  long long to_q;                    /* [512, Dv] */
  long long to_k, to_v;              /* [512, Dv] each, adjacent => one [1024, Dv] operand */
  long long to_out;                  /* [Dv, 512] */
  long long ffw_norm_w, ffw_norm_b;  /* layers.{i}.1.0 */
  long long ffw_w1, ffw_w2;          /* layers.{i}.1.1 [ff_inner,Dv], layers.{i}.1.3 [Dv,ff_inner] */
  long long norm_w, norm_b;          /* final norm */
  long long total;
} fm_resampler_layout;
int fm_resampler_layout_of(const fm_resampler_cfg* cfg, fm_resampler_layout* out);
size_t fm_resampler_saved_bytes(const fm_resampler_cfg* cfg);
size_t fm_resampler_scratch_bytes(const fm_resampler_cfg* cfg);

/* x_f [BN, T, F, Dv]; out [BN*n_latents, Dv] (bf16, or fp32 when out_f32). */
int fm_resampler_fwd(const fm_resampler_cfg* cfg, const float* w_f32, const void* w_bf16, const void* x_f,
                     void* out, int out_f32, void* saved, fm_stream_t stream);
/* dout [BN*n_latents, Dv] bf16; writes every parameter gradient into g_f32 (no gradient for x_f:
 * the CLIP features are produced under no_grad, modeling_flamingo.py:169-170). */
int fm_resampler_bwd(const fm_resampler_cfg* cfg, const float* w_f32, const void* w_bf16, const void* x_f,
                     const void* saved, const void* dout, float* g_f32, void* scratch, fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------ ABI self-check
 * sizeof of the public structs, so a binding can verify its mirror: {gemm_desc, xattn_cfg, xattn_layout,
 * resampler_cfg, resampler_layout}. */
int fm_abi_sizes(int* out5);

/* ------------------------------------------------------------------------------------------------ round-2 entry points
 * (validated on a B200 in round 2: profiles/r02_validate_next/summary.log) */

/* fm_resampler_bwd that reports progress: layer_done(user, l) is called on the calling thread as soon as every kernel
 * writing the gradients of layer l (arena range [layer0 + l*layer_stride, +layer_stride)) has been enqueued on `stream`
 * (side-stream work joined), for l = depth-1 .. 0; a data-parallel caller starts that range's all-reduce right there
 * instead of waiting for the whole arena.  latents / time_pos_emb / final norm are complete when the function returns. */
typedef void (*fm_layer_cb)(void* user, int layer);
int fm_resampler_bwd_notify(const fm_resampler_cfg* cfg, const float* w_f32, const void* w_bf16, const void* x_f,
                            const void* saved, const void* dout, float* grads_f32, void* scratch, fm_layer_cb layer_done,
                            void* user, fm_stream_t stream);
/* The caller's stream waits for everything the library has put on its side stream so far (see FM_OPT_DEFER_JOIN); harmless when
 * nothing is outstanding. */
int fm_side_join(fm_stream_t stream);
int fm_get_option(int key);          /* current value of a switch, -1 for an unknown key */
/* Attention cores on their own (inference): the stand-alone forwards of MaskedCrossAttention (gated_cross_attention.py:
 * 95-124) and PerceiverAttentionLayer (perceiver_resampler.py:79-95) are composed from fm_layernorm_fwd + fm_gemm_bf16 +
 * these.  With I = 64*heads: q: bf16 [rows, I] already scaled by dim_head^-0.5, head h in columns [64h, 64h+64);
 * kv: bf16 [keys, 2I], K in [0,I), V in [I,2I); o: bf16 [rows, I].  dim_head = 64, 64 latents per image, 1..64 heads.
 *   xattn: rows = B*S, keys = B*n_media*64, text_time int32 [B,S] (fm_text_time); masking as in fm_xattn_fwd.
 *   resampler: rows = BN*64 latent queries, keys = BN*nk; lse (optional, fp32 [BN*heads*64]) receives the row log-sum-exp. */
int fm_xattn_core_fwd(const void* q, const void* kv, const int* text_time, void* o, int B, int S, int n_media, int heads,
                      fm_stream_t stream);
int fm_resampler_core_fwd(const void* q, const void* kv, void* o, float* lse, int BN, int nk, int heads, fm_stream_t stream);
/* ... and their backward (stand-alone modules with gradients): d_o = gradient w.r.t. o; dq = q_scale * dS K (gradient w.r.t. the
 * un-scaled query projection); dkv = [dK | dV] in the layout of kv; the resampler one needs the forward's o and lse. */
int fm_xattn_core_bwd(const void* q, const void* kv, const int* text_time, const void* d_o, void* dq, void* dkv, int B, int S,
                      int n_media, int heads, float q_scale, fm_stream_t stream);
int fm_resampler_core_bwd(const void* q, const void* kv, const void* o, const void* d_o, const float* lse, void* dq, void* dkv,
                          int BN, int nk, int heads, float q_scale, fm_stream_t stream);
/* Fused AdamW step over one flat fp32 parameter arena (what the reference gets from HF Trainer's adamw_torch, training/train.sh:10-13):
 * p, g, m, v: fp32 [n] (parameters, gradients, first / second moments), updated in place; shadow_bf16 (optional): the bf16
 * tensor-core copy of the NEW parameters, written in the same pass; decay_mask (optional fp32 [n], 1 = decayed, 0 = not);
 * grad_scale (optional DEVICE scalar multiplied into g: gradient clipping without another pass); step >= 1 (bias correction). */
int fm_adamw_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, const float* decay_mask,
                  const float* grad_scale, long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int step, fm_stream_t stream);
/* Loss head (modeling_flamingo.py:287-298: cross-entropy of logits[..., :-1, :] against labels[..., 1:]), SURVEY §8(f)-3.
 * logits: bf16 [rows, ld], ld a multiple of 8, columns [vocab, ld) are padding (never read; their gradient is zero).
 * targets: int64 [rows]; rows whose target equals ignore_index contribute neither loss nor gradient.
 * fwd: lse[row] = log sum exp(logits[row, :vocab]); row_loss[row] = lse - logits[row, target] (0 if ignored).
 * bwd: dlogits[row, c] = (exp(logits[row, c] - lse[row]) - [c == target]) * (*scale), *scale a DEVICE float
 *      (d loss / number of counted rows), so the whole step stays capturable in a CUDA graph. */
int fm_cross_entropy_fwd(const void* logits, long long ld, int rows, int vocab, const long long* targets,
                         long long ignore_index, float* lse, float* row_loss, fm_stream_t stream);
int fm_cross_entropy_bwd(const void* logits, long long ld, int rows, int vocab, const long long* targets,
                         long long ignore_index, const float* lse, const float* scale, void* dlogits,
                         fm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FLAMINGO_B200_H_ */
