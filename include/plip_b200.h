/*
 * plip_b200 — C ABI of the B200-native PLIP (CLIP ViT-B/32) inference engine.
 *
 * One shared library (libplip_b200.so, nvcc -gencode arch=compute_100a,code=sm_100a) exports
 * exactly the entry points below.  Plain pointers and sizes only: no torch / python types.
 *
 * The reference (PathologyFoundation/plip) has no FFI of its own — its hot path is the Python
 * call surface of a HuggingFace CLIPModel.  Each entry point therefore cites the reference /
 * transformers call it replaces ("TF:" = transformers/models/clip/modeling_clip.py, v5.5.0).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; plip_last_error() then holds a
 *     message (thread-local).  Nothing throws across the ABI.
 *   - *_dev pointers are device pointers owned by the caller; the engine owns packed weights and
 *     its workspace (allocated at plip_create, nothing is allocated on the hot path).
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, the device-pointer
 *     entry points never synchronise the host.  One handle per device; calls on one handle share a
 *     workspace and are serialised on the device (each call waits, stream-side, for the previous one);
 *     a handle must not be used from two host threads at once.
 *   - device-pointer entry points launch on the CALLER'S current device: make the handle's device (or, for the
 *     handle-free functions, the device that owns the buffers) current first, as for any CUDA library call.  The
 *     *_host entry points select the handle's device themselves.
 */
#ifndef PLIP_B200_H_
#define PLIP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLIP_API __attribute__((visibility("default")))

#define PLIP_B200_ABI_VERSION 5  /* 3: + plip_resize_crop_u8; 4: + plip_profile_*, plip_create_ex (operand format); 5: + plip_set_last_layer_pruning */

/* Model constants (TF:configuration_clip.py:47-64,97-109,160-161). */
#define PLIP_IMAGE_SIZE 224
#define PLIP_EMBED_DIM 512
#define PLIP_TEXT_SEQ 77
#define PLIP_VOCAB 49408

typedef struct plip_engine plip_engine_t;

/* ---- pixel formats accepted by plip_encode_images ------------------------------------------ */
enum plip_pixel_format {
  PLIP_PIX_F32_NCHW = 0, /* CLIPProcessor output: normalised float32 [n,3,224,224] (plip.py:35,49) */
  PLIP_PIX_BF16_NCHW = 1, /* same, bfloat16 */
  PLIP_PIX_U8_NHWC = 2   /* raw RGB tiles uint8 [n,224,224,3]; (x/255-mean)/std fused on device
                            (TF:image_processing_clip.py:50-62, embedders/transform.py:51) */
};

enum plip_id_dtype { PLIP_IDS_I32 = 0, PLIP_IDS_I64 = 1 };

/* 16-bit format of every GEMM / attention OPERAND (packed weights, activations between kernels).  Accumulation, the
 * residual stream, LayerNorm statistics, softmax and the similarity head are float32 in both.  tcgen05 kind::f16
 * runs both at the same rate.  BF16 is the default (BASELINE.json's dtype).  FP16 keeps 3 more significand bits:
 * end-to-end |dlogits_per_image| is 6-8x smaller (profiles/r2_precision_study.md) — the reference's own OpenAI-clip
 * flavour runs fp16 weights on the GPU (scripts/extract_embedding.py:94-97) — at the price of a 65504 range. */
enum plip_operand_format { PLIP_OPERAND_BF16 = 0, PLIP_OPERAND_FP16 = 1 };

/* ---- errors / version ---------------------------------------------------------------------- */
PLIP_API const char* plip_last_error(void);
PLIP_API int plip_abi_version(void);
/* Kernel launches issued by the library so far (bench.py's gpu_launches accounting). */
PLIP_API uint64_t plip_launch_count(void);

/* ---- packed weights ------------------------------------------------------------------------ */
/* The engine consumes ONE contiguous host blob.  Its layout is defined by the library and
 * queried by the host-side packer (plip_b200/weights.py), which fills it from a HuggingFace
 * CLIPModel state dict (names listed in SURVEY.md §8a) or an OpenAI-clip state dict.
 * Replaces CLIPModel.from_pretrained / load_state_dict (plip.py:26, embedders/factory.py:21-25). */
typedef struct plip_tensor_info {
  char name[96];    /* HuggingFace-style name, e.g. "vision_model.encoder.layers.0.mlp.fc1.weight" */
  uint64_t offset;  /* byte offset inside the blob (256-byte aligned) */
  uint64_t numel;
  int32_t dtype;    /* 0 = float32, 1 = 16-bit operand (bfloat16, or IEEE half for a PLIP_OPERAND_FP16 engine) */
  int32_t rows;     /* logical 2-D shape (rows x cols), cols == 1 for vectors */
  int32_t cols;
  int32_t fused;    /* 0 = plain copy of the named tensor.
                       bit 0 (1): q/k/v rows concatenated — name holds the q tensor, k and v follow
                         ("...q_proj" -> k_proj, v_proj); the q rows are pre-scaled by head_dim^-0.5 = 0.125.
                       bit 1 (2): the preceding LayerNorm (layer_norm1 for q_proj, layer_norm2 for fc1;
                         TF:371,380) is folded in: weight' = bf16(gamma o W), bias' = bias + W beta, and the
                         pseudo-tensor "<...>.colsum"[n] = sum_k float(weight'[n,k]); the kernels then compute
                         rstd_r * (x_bf16 . weight'_n - mean_r * colsum_n) + bias'_n  ==  LN(x) . W_n + bias_n. */
} plip_tensor_info_t;

PLIP_API int plip_weights_num_tensors(void);
PLIP_API int plip_weights_tensor_info(int index, plip_tensor_info_t* info);
PLIP_API uint64_t plip_weights_blob_bytes(void);

/* ---- engine lifetime ----------------------------------------------------------------------- */
/* host_blob: packed weights (layout above), plus exp(logit_scale) passed separately.
 * max_micro_batch: largest number of images / captions processed per internal pass; larger calls
 * are looped in micro-batches.  Device memory: blob + plip_workspace_bytes(max_micro_batch). */
PLIP_API int plip_create(const void* host_blob, uint64_t blob_bytes, float logit_scale_exp, int device,
                         int max_micro_batch, plip_engine_t** out);
/* Same with an explicit operand format: the 16-bit entries of host_blob must have been packed in that format. */
PLIP_API int plip_create_ex(const void* host_blob, uint64_t blob_bytes, float logit_scale_exp, int device,
                            int max_micro_batch, int operand_format, plip_engine_t** out);
PLIP_API int plip_destroy(plip_engine_t* e);
PLIP_API uint64_t plip_workspace_bytes(int max_micro_batch);
PLIP_API float plip_logit_scale_exp(const plip_engine_t* e);
PLIP_API int plip_max_micro_batch(const plip_engine_t* e);
PLIP_API int plip_operand_format(const plip_engine_t* e);
/* Pooled position of a caption WITHOUT an eos token (49407).  0 (default): position 0, as transformers does for
 * configs with eos_token_id == 49407 (TF:571-584: argmax of an all-zero match vector).  1: first position of the
 * largest id, as legacy configs (eos_token_id == 2 — what openai/clip-vit-base-patch32 ships) and OpenAI clip's
 * text.argmax(-1) do (TF:564-570).  Captions that contain an eos are pooled at its first position either way. */
PLIP_API int plip_set_text_pooling(plip_engine_t* e, int no_eos_argmax);
/* Last-layer pruning for the embedding calls (default 0 = off).  Only the pooled row of a sequence leaves a tower
 * (CLS: TF:685-686; first eos: TF:571-584), and after the last layer's attention rows no longer interact, so with
 * on != 0 plip_encode_* / plip_clip_forward* run that layer's out_proj, LayerNorm 2 and MLP on the n pooled rows
 * instead of all n*seq rows: the embeddings are the same (same per-row arithmetic), the step is ~5 % shorter at
 * batch 1024 and more for small batches.  plip_hidden_states is never pruned.  plip_last_layer_pruning reads it back. */
PLIP_API int plip_set_last_layer_pruning(plip_engine_t* e, int on);
PLIP_API int plip_last_layer_pruning(const plip_engine_t* e);

/* ---- the hot path -------------------------------------------------------------------------- */
/* Vision tower + visual_projection: replaces CLIPModel.get_image_features (TF:829-863, called at
 * plip.py:50) and OpenAI-clip model.encode_image (embedders/plip.py:48).
 * out_dev: float32 [n,512]; normalize != 0 divides each row by its L2 norm (TF:57-65,923). */
PLIP_API int plip_encode_images(plip_engine_t* e, const void* pixels_dev, int pixel_format, int64_t n,
                                float* out_dev, int normalize, void* stream);

/* Text tower + text_projection: replaces CLIPModel.get_text_features (TF:793-825, called at
 * plip.py:68) and model.encode_text (embedders/plip.py:66).
 * ids_dev: [n,seq_len] token ids (seq_len <= 77); attention_mask_dev: optional [n,seq_len] of the
 * same dtype (0 = padded key), NULL = no padding mask.  Pooling takes the row of the first
 * eos_token_id (49407), as TF:571-584. */
PLIP_API int plip_encode_text(plip_engine_t* e, const void* ids_dev, int ids_dtype,
                              const void* attention_mask_dev, int64_t n, int seq_len, float* out_dev,
                              int normalize, void* stream);

/* Same, processing only the first prefix_len (<= seq_len) positions of every row.  The pooled output of a
 * caption depends only on positions up to its first eos (causal attention, TF:546-557,571-584), so with
 * prefix_len >= (longest first-eos position + 1) the result equals plip_encode_text at prefix_len/seq_len of
 * the work — what typical prompts ("An H&E image patch of ...", ~12 of 77 tokens) need. */
PLIP_API int plip_encode_text_prefix(plip_engine_t* e, const void* ids_dev, int ids_dtype,
                                     const void* attention_mask_dev, int64_t n, int seq_len, int prefix_len,
                                     float* out_dev, int normalize, void* stream);

/* Similarity head: logits_per_image[n,m] = scale * norm(img)[n,512] . norm(txt)[m,512]^T
 * (TF:923-930; numpy versions at plip.py:73-76, evaluation/zero_shot/zero_shot.py:12,
 * evaluation/retrieval/retrieval.py:14).  normalize_img / normalize_txt select which side is
 * L2-normalised on the fly (PLIP._cosine_similarity normalises only the key side). */
PLIP_API int plip_similarity(const float* img_dev, int64_t n, const float* txt_dev, int64_t m, float scale,
                             int normalize_img, int normalize_txt, float* logits_dev, int64_t ld_logits,
                             void* stream);

/* Fused similarity + top-k over the second operand (k <= 64), never materialising [n,m]:
 * idx_dev int32 [n,k] (descending score), val_dev float32 [n,k] (may be NULL).
 * Replaces np.argmax (plip.py:102, zero_shot.py:13) and argsort()[:, -k:][:, ::-1]
 * (plip.py:85, retrieval.py:16). */
PLIP_API int plip_similarity_topk(const float* query_dev, int64_t n, const float* space_dev, int64_t m,
                                  float scale, int normalize_query, int normalize_space, int k,
                                  int32_t* idx_dev, float* val_dev, void* stream);

/* L2-normalise rows in place (embedders/plip.py:53,73). */
PLIP_API int plip_l2_normalize(float* x_dev, int64_t n, int dim, void* stream);

/* Image preparation on the device: n variable-size RGB uint8 images (HWC, rows packed, image i at byte
 * descs[i].offset of src_dev) -> tiles_dev uint8 [n,224,224,3] (the PLIP_PIX_U8_NHWC input of
 * plip_encode_images).  Image i is resized to new_width x new_height with Pillow's antialiased bicubic filter
 * and the 224x224 window at (left, top) of the resized image is kept — bit-identical to
 * PIL.Image.resize((new_width,new_height), BICUBIC).crop(...), i.e. to what CLIPProcessor's resize +
 * center_crop (plip.py:35; TF:models/clip/image_processing_clip.py:50-62) and torchvision's
 * Resize(224, BICUBIC) + CenterCrop(224) (reproducibility/embedders/transform.py:45-52) produce before their
 * float conversion.  descs_host is a HOST array (the caller knows the sizes from decoding); it is consumed
 * before the call returns.  src_dev / tiles_dev are device pointers; stream-ordered, no synchronisation. */
typedef struct plip_resize_desc {
  int64_t offset;                 /* byte offset of the image in src_dev */
  int32_t width, height;          /* source size */
  int32_t new_width, new_height;  /* size after the resize (each >= 224) */
  int32_t left, top;              /* crop origin in the resized image */
} plip_resize_desc_t;
PLIP_API int plip_resize_crop_u8(const void* src_dev, uint64_t src_bytes, const plip_resize_desc_t* descs_host,
                                 int64_t n, void* tiles_dev, void* stream);

/* ---- host-buffer convenience (end-to-end path; copies are inside the call) ------------------- */
/* pixels_host / ids_host / out_host are host pointers (pinned or pageable).  The call stages
 * micro-batches through pinned buffers on two streams (H2D overlapped with compute), writes the
 * float32 [n,512] result to out_host and returns after the last D2H completed.
 * plip_encode_text_host scans the ids for the caption lengths (first eos) and uses plip_encode_text_prefix: one
 * pass up to the longest caption, or — for large batches of mixed lengths — the captions sorted by length and
 * processed in up to 8 length buckets, each only up to its own longest caption (results are returned in the
 * caller's order). */
PLIP_API int plip_encode_images_host(plip_engine_t* e, const void* pixels_host, int pixel_format, int64_t n,
                                     float* out_host, int normalize);
PLIP_API int plip_encode_text_host(plip_engine_t* e, const void* ids_host, int ids_dtype,
                                   const void* attention_mask_host, int64_t n, int seq_len,
                                   float* out_host, int normalize);

/* ---- in-step kernel timing (measurement support for bench.py) --------------------- */
/* While enabled, every kernel launch of plip_encode_images / plip_encode_text* on this handle is bracketed by a
 * CUDA event pair recorded on the launch stream.  plip_profile_read waits for the recorded events and returns one
 * aggregated row per (tower, kernel role): launches, summed device time, and the ALGORITHMIC flops / HBM bytes of
 * those launches — i.e. each kernel's average duration inside the step it belongs to, under the
 * step's own clocks and cache state.  Event pairs serialise nothing but cost a few microseconds of launch gap
 * each, so bench.py profiles separate, untimed steps.  enable(on) always clears what was recorded. */
typedef struct plip_kernel_time {
  char name[48];     /* "<tower>/<role>", e.g. "vision/gemm[fc2+resid]", "text/attention" */
  int32_t launches;
  float total_ms;
  double flops;      /* algorithmic FLOPs of the recorded launches (0 for memory-bound helpers) */
  double bytes;      /* algorithmic HBM bytes of the recorded launches */
} plip_kernel_time_t;
PLIP_API int plip_profile_enable(plip_engine_t* e, int on);
PLIP_API int plip_profile_read(plip_engine_t* e, plip_kernel_time_t* out, int cap, int* count);

/* ---- per-kernel test hooks (used by tests/ only; stream-ordered, device pointers) ------------ */
/* The handle-free hooks below interpret / produce 16-bit data in the format set here (default PLIP_OPERAND_BF16). */
PLIP_API int plip_dbg_set_operand_format(int operand_format);
/* epilogue: 0 bias->bf16, 1 bias+QuickGELU->bf16, 2 x_f32 += acc+bias (optionally also xb_out bf16 copy +
 * stats_out [M,8,2] row statistics), 3 patch scatter + pos, 4 plain f32, 5/6 = 0/1 with the LayerNorm fold
 * (colsum [N], stats_in [M,8,2] with n_partials valid slots). */
PLIP_API int plip_dbg_gemm(const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K,
                           const float* bias, void* out, int ldo, const float* pos, int epilogue, int cta_group,
                           int block_n, const float* colsum, const float* stats_in, int n_partials, void* xb_out,
                           float* stats_out, void* stream);
/* Host-only: fixed-point filter row of output index xx for one axis of plip_resize_crop_u8 (the same code the
 * kernel runs).  Returns the filter bank width ksize (or -ksize if k_cap is too small). */
PLIP_API int plip_dbg_resize_filter(int in_size, int out_size, int xx, int32_t* k_host, int k_cap, int* xmin,
                                    int* count);
/* Host-only: the length-bucket plan plip_encode_text_host uses (lens = first-eos position + 1 per caption).
 * perm_host[i] = original index of the caption at sorted position i (may be NULL); bucket k covers sorted positions
 * [bucket_start[k], bucket_start[k+1]) and is processed with prefix bucket_prefix[k].  Arrays hold cap+1 / cap
 * entries (cap <= 8).  Returns the number of buckets. */
PLIP_API int plip_dbg_text_bucket_plan(const int32_t* lens_host, int64_t n, int seq_len, int32_t* perm_host,
                                       int32_t* bucket_start_host, int32_t* bucket_prefix_host, int cap);
PLIP_API int plip_dbg_rowstats_cast(const float* x, int64_t rows, int dim, void* xb_bf16, float* stats, void* stream);
PLIP_API int plip_dbg_layernorm(const float* x, int64_t rows, int dim, int64_t in_row_stride,
                                const float* gamma, const float* beta, float* out_f32, void* out_bf16,
                                void* stream);
PLIP_API int plip_dbg_attention(const void* qkv_bf16, int64_t n_seq, int seq_len, int heads, int causal,
                                const int32_t* key_mask, void* out_bf16, void* stream);
PLIP_API int plip_dbg_im2col(const void* pixels, int pixel_format, int64_t n, void* out_bf16, void* stream);
/* Run one tower and copy the fp32 residual stream [n*S, D] after `num_layers` encoder layers
 * (0 = after embeddings / pre-LN) into hidden_dev.  tower: 0 = vision (input pixels), 1 = text. */
PLIP_API int plip_dbg_hidden_states(plip_engine_t* e, int tower, const void* input_dev, int input_format,
                                    const void* attention_mask_dev, int64_t n, int num_layers,
                                    float* hidden_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLIP_B200_H_ */
