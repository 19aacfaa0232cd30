/*
 * pvsg.h -- C ABI of libpvsg_sm100.so, the B200 (sm_100a) backend of the OpenPVSG
 * inference hot path (Mask2Former / Mask2Former-VPS forward + relation head).
 *
 * Conventions:
 *   - plain device pointers + explicit sizes; no torch types; fp32 unless stated;
 *   - feature maps are TOKEN-MAJOR (NHWC): [B, H, W, C] == [B*H*W, C], C contiguous;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     allocates, never synchronises, keeps no global state -> CUDA-graph capturable;
 *   - returns PVSG_OK (0) or a negative error code (see pvsg_error_string); a failed
 *     argument check launches nothing.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * OpenPVSG tree; "L0" = inside mmcv-full 1.4.0 / mmdet 2.25.0, reached from the cited
 * reference call site).
 */
#ifndef PVSG_H_
#define PVSG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVSG_VERSION 100
#define PVSG_OK 0
#define PVSG_ERR_INVALID_ARG (-1)
#define PVSG_ERR_UNSUPPORTED (-2)
#define PVSG_ERR_LAUNCH (-3)
#define PVSG_ERR_NO_DEVICE (-4)

#define PVSG_ACT_NONE 0
#define PVSG_ACT_RELU 1
#define PVSG_ACT_GELU 2 /* exact erf GELU (torch nn.GELU default; Swin FFN) */

int pvsg_version(void);
const char* pvsg_error_string(int code);
/* SM count / compute capability of `device`; PVSG_ERR_NO_DEVICE when no CUDA device. */
int pvsg_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* Handle ("handle + workspace").  The op entry points below are stateless and need no handle; a handle is
 * the create-time half of the contract for a host that wants it:
 *   pvsg_create   checks that `device` is compute capability 10.x (PVSG_ERR_UNSUPPORTED otherwise), and sets every
 *                 kernel's opt-in attributes (dynamic shared memory, carve-out) on it, so that no op call made later --
 *                 for instance the first one, inside a stream capture -- has a first-use side effect;
 *   pvsg_workspace returns a device scratch pointer of at least `bytes` owned by the handle (grow-only; growing frees
 *                 the old block, i.e. synchronises -- size it before capturing), for the `ws` arguments sized by the
 *                 pvsg_<op>_workspace_bytes companions (attention, attention_tc, attention_t5, reconsdot);
 *   pvsg_destroy  frees it.  Replaces nothing in the reference (torch's caching allocator plays this part there). */
typedef struct pvsg_handle pvsg_handle;
int pvsg_create(int device, pvsg_handle** out);
int pvsg_destroy(pvsg_handle* handle);
int pvsg_handle_info(const pvsg_handle* handle, int* device, int* sm_count, int64_t* smem_optin_bytes, int64_t* workspace_bytes);
int pvsg_workspace(pvsg_handle* handle, int64_t bytes, void** ptr);

/* ---------------------------------------------------------------- dense layers --- */

/* C[b] = act(A[b] (+A2[b]) . W[b]^T + bias + R[b]);  A [M,K] (row stride lda), W [N,K]
 * (row stride ldw), C [M,N] (ldc), R [M,N] (ldr) optional, A2 same layout as A, optional.
 * `batch` problems with element strides sA/sW/sC (sA also applies to A2, sC to R).
 * Replaces torch nn.Linear / F.linear at: mask2former_head.py:376-380 (cls_embed,
 * mask_embed), L0 FFN / value_proj / sampling_offsets / attention_weights / output_proj /
 * in_proj / out_proj (configs/mask2former_vps/mask2former_video_r50_base.py:36-88), and
 * models/relation_head/{base.py:46-48, transformer.py:28-32}. */
int pvsg_linear(const float* A, const float* A2, const float* W, const float* bias,
                const float* R, float* C, int64_t M, int64_t N, int64_t K,
                int64_t lda, int64_t ldw, int64_t ldc, int64_t ldr, int act,
                int64_t batch, int64_t sA, int64_t sW, int64_t sC, void* stream);

/* y = act(conv2d(x, w) + bias + residual), NHWC, implicit GEMM (no im2col buffer).
 * x [B,H,W,Cin]; w [Cout,R,S,Cin] (BN already folded in by the host); y/residual
 * [B,OH,OW,Cout], OH = (H + 2*pad - R)/stride + 1.
 * Replaces the cuDNN convs of L0 ResNet / ConvModule reached from
 * models/mask2former_vps/mask2former.py:130 and mask2former_video_head.py:382. */
int pvsg_conv2d_nhwc(const float* x, const float* w, const float* bias, const float* residual,
                     float* y, int B, int H, int W, int Cin, int Cout, int R, int S,
                     int stride, int pad, int act, void* stream);

/* ---- tcgen05 engine (same contractions on the 5th-gen tensor cores, fp32-grade) ----
 * Operands are "split-bf16": an fp32 tensor x is carried as two bf16 planes of the same
 * shape, hi = bf16(x), lo = bf16(x - hi) (4 bytes / element, like fp32); one product is
 * hi.hi + hi.lo + lo.hi accumulated in fp32 in TMEM (error ~2^-17 relative per product).
 * TMA (cp.async.bulk.tensor, 128B swizzle) feeds tcgen05.mma kind::f16 from shared memory. */

/* hi/lo planes of x (+ x2 if non-null); n elements, n % 4 == 0. */
int pvsg_split_bf16(const float* x, const float* x2, void* hi, void* lo, int64_t n, void* stream);

/* pvsg_linear on split operands.  A [M,K] planes (row stride lda), W [N,K] planes (ldw);
 * outputs (any subset): C fp32, (C_hi, C_lo) split planes, or mask/row_open = the sign-mask
 * epilogue of pvsg_mask_logits; all with row stride ldc.  K % 64 == 0, lda/ldw % 8 == 0.
 * The residual is either fp32 (R) or itself split planes (R_hi, R_lo; r = hi + lo), row stride
 * ldr -- activations can then live as planes only (no fp32 copy is written or read).  With
 * `ident` (a bf16 [256,256] identity matrix in device memory, supplied by the caller because the
 * library never allocates) the plane residual is added BY THE TENSOR CORE as extra k-blocks
 * R_hi.I + R_lo.I, exact in the fp32 accumulator, so its bytes ride the main TMA pipeline.
 * Returns PVSG_ERR_UNSUPPORTED for shapes it does not cover (caller uses pvsg_linear). */
int pvsg_linear_tc(const void* A_hi, const void* A_lo, int64_t lda, const void* W_hi,
                   const void* W_lo, int64_t ldw, const float* bias, const float* R, int64_t ldr,
                   float* C, void* C_hi, void* C_lo, uint8_t* mask, int32_t* row_open, int64_t ldc,
                   int64_t M, int64_t N, int64_t K, int act, const void* R_hi, const void* R_lo,
                   const void* ident, void* stream);

/* `batch` independent contractions C[b] = A[b] . W[b]^T in ONE launch (3-D tensor maps; batch strides
 * a_bs / w_bs in elements): the per-frame query x pixel mask-logit einsum of a batch of frames
 * ('bqc,bchw->bqhw', mask2former_head.py:382).  Outputs as pvsg_mask_logits: C fp32 [batch,M,ldc] and/or
 * mask uint8 [batch,M,ldc] (1 where the logit < 0) + row_open int32 [batch,M] (+=, caller zeroes). */
int pvsg_linear_tc_batched(const void* A_hi, const void* A_lo, int64_t lda, int64_t a_bs,
                           const void* W_hi, const void* W_lo, int64_t ldw, int64_t w_bs, float* C,
                           uint8_t* mask, int32_t* row_open, int64_t ldc, int batch, int64_t M,
                           int64_t N, int64_t K, void* stream);

/* pvsg_conv2d_nhwc (stride 1 or 2) on split operands: x planes [B,H,W,Cin], w planes
 * [Cout,R,S,Cin]; im2col-free -- a 4-D TMA box over the NHWC planes is shifted per filter
 * tap (traversed with elementStrides = stride) and out-of-bounds zero fill provides the
 * padding.  Cin % 64 == 0.  residual fp32 or split planes (res_hi, res_lo) [B,OH,OW,Cout]. */
int pvsg_conv2d_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                   const float* bias, const float* residual, float* y, void* y_hi, void* y_lo,
                   int B, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad,
                   int act, const void* res_hi, const void* res_lo, void* stream);

/* ResNet stem (7x7, stride 2, pad 3, C <= 4 input channels; mmdet ResNet conv1) without im2col:
 * packs x [B,C,H,W] (NCHW fp32, the detector's input layout) into operand planes
 * X2 [B, OH+3, OW, 64], X2[b,yp,ox, par*32 + s*4 + c] = x[b, c, 2yp+par-3, 2ox+s-3] (zero outside
 * the frame and for s = 7 / c >= C), OH = (H-1)/2+1, OW = (W-1)/2+1.  The stem is then the 4 x 1
 * stride-1 convolution pvsg_conv2d_tc(X2, W2 [Cout,4,1,64]) with W2[co, j, 0, par*32 + s*4 + c] =
 * w[co, 2j+par, s, c]. */
int pvsg_stem7x7s2_pack(const float* x_nchw, void* hi, void* lo, int B, int C, int H, int W,
                        void* stream);

/* Small-Cin convolutions (generic form, Cin % 64 != 0): gathers the patches of x [B,H,W,Cin]
 * directly into split operand planes [B*OH*OW, Kpad] (k = (r*S + s)*Cin + c, zero-padded to
 * Kpad >= R*S*Cin, Kpad % 64 == 0 for pvsg_linear_tc), so the stem also runs on tcgen05. */
int pvsg_im2col_split(const float* x, void* hi, void* lo, int B, int H, int W, int Cin, int R,
                      int S, int stride, int pad, int Kpad, void* stream);

/* 3x3 stride-2 pad-1 max pooling, NHWC (L0 ResNet stem). */
int pvsg_maxpool3x3s2_nhwc(const float* x, float* y, int B, int H, int W, int C, void* stream);

/* layout converters between the reference's NCHW tensors and token-major. */
int pvsg_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, void* stream);
int pvsg_nhwc_to_nchw(const float* x, float* y, int B, int C, int H, int W, void* stream);

/* y[r,:] = LayerNorm(x[r,:]) * gamma + beta, eps inside sqrt; C in {128,256,512,1024}.
 * (torch nn.LayerNorm: L0 norms / post_norm mask2former_head.py:375; relation
 * transformer.py:26,44.) */
int pvsg_layernorm(const float* x, const float* gamma, const float* beta, float* y,
                   int64_t rows, int C, float eps, void* stream);
/* same, additionally emitting the split-bf16 planes (y_hi, y_lo) of y for a following
 * pvsg_linear_tc (saves a pvsg_split_bf16 pass). */
int pvsg_layernorm_split(const float* x, const float* gamma, const float* beta, float* y,
                         void* y_hi, void* y_lo, int64_t rows, int C, float eps, void* stream);

/* same as pvsg_layernorm_split, additionally the planes (s_hi, s_lo) of y + add (add [rows,C]):
 * the (query + query_pos) operand of the next MSDeformAttn's offset / weight projections. */
int pvsg_layernorm_split2(const float* x, const float* gamma, const float* beta, float* y,
                          void* y_hi, void* y_lo, const float* add, void* s_hi, void* s_lo,
                          int64_t rows, int C, float eps, void* stream);

/* GroupNorm over token-major x [B,HW,C] (L0 ConvModule norm GN(32)); stats = workspace
 * of 2*B*groups doubles (zeroed by the call). y = act(GN(x)). */
int pvsg_groupnorm_nhwc(const float* x, const float* gamma, const float* beta, float* y,
                        double* stats, int B, int64_t HW, int C, int groups, float eps,
                        int act, void* stream);

/* same, emitting the split planes of y (y itself optional: may be NULL). */
int pvsg_groupnorm_nhwc_split(const float* x, const float* gamma, const float* beta, float* y,
                              void* y_hi, void* y_lo, double* stats, int B, int64_t HW, int C,
                              int groups, float eps, int act, void* stream);

/* y[r,:] = x[r,:] + v[:]  (level_embed add, mask2former_head.py:424-426). */
int pvsg_add_rowvec(const float* x, const float* v, float* y, int64_t rows, int C, void* stream);

/* Bilinear resize, align_corners=False, token-major: dst [B,OH,OW,C] (= or +=) resize(src
 * [B,IH,IW,C]).  (F.interpolate at L0 pixel-decoder FPN step, and the source-side
 * pooling that the attn-mask downsample of mask2former_head.py:383-387 reduces to.) */
int pvsg_bilinear_resize_nhwc(const float* src, float* dst, int B, int IH, int IW, int OH,
                              int OW, int C, int accumulate, void* stream);
/* same with a source batch stride (elements; frames of a batch that are slices of a longer token
 * buffer) and optional split-bf16 planes (dst_hi, dst_lo) of the final dst values. */
int pvsg_bilinear_resize_nhwc_ex(const float* src, int64_t src_batch_stride, float* dst, void* dst_hi,
                                 void* dst_lo, int B, int IH, int IW, int OH, int OW, int C,
                                 int accumulate, void* stream);

/* Sine positional encoding, token-major out [T*H*W, 2*num_feats] (+ add_vec[2*num_feats]
 * if non-null).  dim_t [num_feats] / dim_t_z [2*num_feats] are the temperature tables
 * (host-computed exactly as the reference does).  T = 0 -> 2-D encoding (L0
 * SinePositionalEncoding); T >= 1 -> models/mask2former_vps/position_encoding.py:55-99. */
int pvsg_sine_pe(float* out, const float* dim_t, const float* dim_t_z, const float* add_vec,
                 int T, int H, int W, int num_feats, float scale, float eps, void* stream);

/* ------------------------------------------------- multi-scale deformable attn --- */

/* mmcv MultiScaleDeformableAttnFunction.forward (L0; selected by cfg
 * mask2former_video_r50_base.py:38-47).  value [B,N,H,D]; spatial_shapes [L,2] (h,w) and
 * level_start_index [L] are HOST int64 arrays; sampling_locations [B,Nq,H,L,P,2] in
 * [0,1] (x,y); attention_weights [B,Nq,H,L,P]; out [B,Nq,H*D].  D must be 32. */
int pvsg_msda_forward(const float* value, const int64_t* spatial_shapes,
                      const int64_t* level_start_index, const float* sampling_locations,
                      const float* attention_weights, float* out, int B, int64_t N,
                      int64_t Nq, int H, int D, int L, int P, void* stream);

/* Fused variant: takes the RAW projections proj [B,Nq,H*L*P*3] = [offsets (H,L,P,2) |
 * attention logits (H,L,P)] and reference points ref [Nq,2] (x,y in [0,1], shared by all
 * levels) and does softmax over L*P, location arithmetic, bilinear sampling and the
 * weighted sum in one kernel (mmcv MultiScaleDeformableAttention.forward minus its
 * Linear layers). */
int pvsg_msda_fused_forward(const float* value, const int64_t* spatial_shapes,
                            const int64_t* level_start_index, const float* proj,
                            const float* ref, float* out, int B, int64_t N, int64_t Nq,
                            int H, int D, int L, int P, void* stream);

/* same, emitting the split planes of the result for the output projection (out optional). L*P <= 16. */
int pvsg_msda_fused_forward_split(const float* value, const int64_t* spatial_shapes,
                                  const int64_t* level_start_index, const float* proj,
                                  const float* ref, float* out, void* out_hi, void* out_lo, int B,
                                  int64_t N, int64_t Nq, int H, int D, int L, int P, void* stream);

/* ------------------------------------------------------------------ attention --- */

/* out = softmax(scale * Q K^T + mask) V per (batch, head); token strides in floats:
 * element (b, i, h, d) of Q is Q[b*q_bs + i*q_ts + h*D + d] (same for K, V, out).
 * mask: optional uint8 [B,Lq,Lk], non-zero = blocked, shared by all heads
 * (mask2former_head.py:388-391 repeats it per head); row_open: optional int32 [B,Lq]
 * = number of un-blocked keys of that row -- rows with 0 attend to everything
 * (mask2former_head.py:453-454).  ws: workspace of pvsg_attention_workspace_bytes.
 * D in {32, 128}.  Replaces nn.MultiheadAttention core (L0 MultiheadAttention reached
 * from mask2former_head.py:457-468) and the encoder self-attention of
 * models/relation_head/base.py:32-37, transformer.py:20-25. */
int64_t pvsg_attention_workspace_bytes(int B, int H, int Lq, int Lk, int D);
int pvsg_attention(const float* Q, const float* K, const float* V, const uint8_t* mask,
                   const int32_t* row_open, float* out, void* ws, int B, int H, int Lq,
                   int Lk, int D, int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts,
                   int64_t v_bs, int64_t v_ts, int64_t o_bs, int64_t o_ts, float scale,
                   void* stream);

/* Same contraction on the tensor cores for head dim 32 or 128 (the decoder's masked cross-attention,
 * mask2former_head.py:457-468; the relation head's encoders, base.py:32-37, transformer.py:20-25): K and V are given as the split-bf16 operand planes their
 * projection emitted (element (b, i, h, d) at b*k_bs + i*k_ts + h*D + d, strides in elements),
 * Q / out fp32 as above; every product is three m16n8k16 MMAs (fp32-grade).  Workspace:
 * pvsg_attention_tc_workspace_bytes. */
int64_t pvsg_attention_tc_workspace_bytes(int B, int H, int Lq, int Lk, int D);
int pvsg_attention_tc(const float* Q, const void* K_hi, const void* K_lo, const void* V_hi,
                      const void* V_lo, const uint8_t* mask, const int32_t* row_open, float* out,
                      void* ws, int B, int H, int Lq, int Lk, int D, int64_t q_bs, int64_t q_ts,
                      int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts, int64_t o_bs,
                      int64_t o_ts, float scale, void* stream);

/* ------------------------------------------------------------ mask logits ------- */

/* The per-frame query x pixel contraction einsum('bqc,bchw->bqhw')
 * (mask2former_head.py:382, mask2former_video_head.py:344) on token-major features:
 * logits[b,q,p] = sum_c embed[b,q,c] * feat[b,p,c].
 *   logits    : optional fp32 [B,Q,P] output;
 *   attn_mask : optional uint8 [B,Q,P] = (logit < 0)  -- i.e. sigmoid < 0.5,
 *               mask2former_head.py:391 -- with row_open[b,q] = #(logit >= 0)
 *               (int32, zeroed by the call).
 * For the attention mask of the NEXT decoder layer pass feat = the mask features
 * bilinearly resized to that layer's (h_l, w_l) (pvsg_bilinear_resize_nhwc): the
 * align_corners=False downsample is linear, so it commutes with the contraction. */
int pvsg_mask_logits(const float* embed, const float* feat, float* logits, uint8_t* attn_mask,
                     int32_t* row_open, int B, int Q, int64_t P, int C, void* stream);

/* --------------------------------------------------------- panoptic fusion ------ */

/* Fused replacement of: final x4 bilinear upsample (mask2former_head.py:674-679 /
 * mask2former_video_head.py:662-667) + crop + optional rescale + sigmoid + score-weighted
 * argmax + per-segment area tests + id assignment
 * (mask2former_fusion_head.py:96-171, :373-383).  The upsampled logits never exist in
 * memory and there is no host round trip.
 *   cls_logits [Q,NC+1]; mask_logits [Q,h,w] (low-res, un-upsampled);
 *   (in_h,in_w) = batch_input_shape, (img_h,img_w) = img_shape crop, (out_h,out_w) =
 *   ori_shape if rescale else img_shape;
 *   pan_out  int32 [out_h,out_w] (void = NC);
 *   seg_info int32 [1 + 4*Q]: [0] = n_kept, then per kept query k (query order):
 *            (query index, class, seg_id or -1 if dropped, final area);
 *   work     int32 [4*Q] scratch (areas); scores float [Q] scratch;
 *   pix_ws   uint16 [out_h*out_w] scratch (winner index + ">= 0.5" bit per pixel). */
int pvsg_panoptic_fuse(const float* cls_logits, const float* mask_logits, int Q, int NC,
                       int num_things, int h, int w, int in_h, int in_w, int img_h, int img_w,
                       int out_h, int out_w, float object_mask_thr, double iou_thr,
                       int filter_low_score, int instance_offset, int32_t* pan_out,
                       int32_t* seg_info, int32_t* work, float* scores, uint16_t* pix_ws,
                       void* stream);

/* Instance branch (mask2former_fusion_head.py:192-242) for `n` candidate queries already
 * selected by the caller's top-k: for each candidate the binary mask (upsampled logit > 0)
 * statistics at output resolution: stats float [n,2] = (sum of sigmoid over the mask, pixel
 * count); boxes int32 [n,4] = mask2bbox (x_min, y_min, x_max+1, y_max+1), zeros when empty;
 * masks_out (optional) uint8 [n,out_h,out_w]. */
int pvsg_instance_masks(const float* mask_logits, const int32_t* query_idx, int n, int h, int w,
                        int in_h, int in_w, int img_h, int img_w, int out_h, int out_w,
                        float* stats, int32_t* boxes, uint8_t* masks_out, void* stream);

/* Instance candidates (mask2former_fusion_head.py:214-222): scores = softmax(cls_logits)[:, :NC]
 * (cls_logits [Q,NC+1], last column = void), then the k largest entries of the flattened
 * [Q*NC] scores -- scores.flatten(0, 1).topk(k, sorted=False).  Outputs [k]: score, label
 * (= index % NC) and query (= index / NC), in ascending flat-index order (torch leaves the order of
 * sorted=False unspecified). */
int pvsg_instance_select(const float* cls_logits, int Q, int NC, int k, float* top_scores,
                         int32_t* top_labels, int32_t* top_query, void* stream);

/* The detector's instance selection (models/mask2former_vps/mask2former.py:192-201) with static
 * shapes: for the n candidates of pvsg_instance_select / pvsg_instance_masks, det_score = score *
 * stats[:,0] / (stats[:,1] + 1e-6) for thing labels (< num_things), stuff candidates rank last;
 * sorted descending, the best `topk` are written: boxes6 [topk,6] = (1-based id among the thing
 * candidates, x0, y0, x1, y1, det_score), labels [topk], sel_query [topk]; count[0] = number of
 * thing candidates (rows beyond it are padding the caller drops). */
int pvsg_instance_finalize(const float* scores, const int32_t* labels, const int32_t* query,
                           const float* stats, const int32_t* boxes, int n, int num_things, int topk,
                           float* boxes6, int32_t* out_labels, int32_t* sel_query, int32_t* count,
                           void* stream);

/* Batched forms (B frames per launch; every array gains a leading [B] axis, mask_logits is
 * [B,Q,h,w], query_idx [B,n] indexes the queries of its own frame).  The single-frame entry
 * points above are these with B = 1. */
int pvsg_panoptic_fuse_batched(const float* cls_logits, const float* mask_logits, int B, int Q, int NC,
                               int num_things, int h, int w, int in_h, int in_w, int img_h, int img_w,
                               int out_h, int out_w, float object_mask_thr, double iou_thr,
                               int filter_low_score, int instance_offset, int32_t* pan_out,
                               int32_t* seg_info, int32_t* work, float* scores, uint16_t* pix_ws,
                               void* stream);
int pvsg_instance_select_batched(const float* cls_logits, int B, int Q, int NC, int k,
                                 float* top_scores, int32_t* top_labels, int32_t* top_query,
                                 void* stream);
int pvsg_instance_masks_batched(const float* mask_logits, const int32_t* query_idx, int B, int Q, int n,
                                int h, int w, int in_h, int in_w, int img_h, int img_w, int out_h,
                                int out_w, float* stats, int32_t* boxes, uint8_t* masks_out,
                                void* stream);
int pvsg_instance_finalize_batched(const float* scores, const int32_t* labels, const int32_t* query,
                                   const float* stats, const int32_t* boxes, int B, int n,
                                   int num_things, int topk, float* boxes6, int32_t* out_labels,
                                   int32_t* sel_query, int32_t* count, void* stream);

/* Run-length events of the kept segments of B panoptic maps for the tube wire format (reference:
 * concat_seq, models/mask2former_vps/utils.py:38-54, pycocotools RLE of `pan == id`, masks.txt rows of
 * models/unitrack/utils/io.py:14-37).  pan int32 [B,H,W]; seg_info [B,1+4Q] as written by
 * pvsg_panoptic_fuse (segment slot = order of first appearance of a segment id among the kept
 * rows).  Outputs per frame, in COLUMN-major walk order: ev_pos uint32 [B,cap] (position x*H + y
 * where a run of the segment starts or ends), ev_slot int16 [B,cap], n_events int32 [B] (may
 * exceed cap: the caller then falls back to the map).  RLE counts of a segment = differences of
 * its positions, with 0 prepended and H*W appended.  col_ws: int32 workspace [2,B,W]. */
int pvsg_rle_events(const int32_t* pan, const int32_t* seg_info, int B, int Q, int H, int W,
                    int32_t* col_ws, uint32_t* ev_pos, int16_t* ev_slot, int32_t* n_events, int cap,
                    void* stream);

/* HOST function (CPU): the events of one frame (after the D2H copy) -> pycocotools RLE strings.
 * Stable counting sort by slot, run lengths = position differences (0 prepended, hw = H*W
 * appended), rleToString encoding.  out: byte buffer of out_cap bytes; seg_off int64 [nseg+1]:
 * string k = out[seg_off[k] .. seg_off[k+1]).  Returns the total length or a negative error. */
int64_t pvsg_rle_strings_host(const uint32_t* ev_pos, const int16_t* ev_slot, int64_t n, int nseg,
                              uint32_t hw, char* out, int64_t out_cap, int64_t* seg_off);

/* ------------------------------------------------------------ Swin backbone ----- */
/* BASELINE configs[2] names a Swin-B backbone; the reference ships none, so the interface replaced
 * here is mmdet 2.25.0's (the version the reference pins, README.md:123-125):
 * mmdet/models/backbones/swin.py ShiftWindowMSA.forward :175-246 + WindowMSA.forward :84-117
 * between the qkv and proj linears.  qkv fp32 [B,H,W,3C] (channel = which*C + head*32 + d) as the
 * qkv linear wrote it for the UNPADDED map; qkv_bias [3C] stands in for the zero-padded positions;
 * bias_table [(2*window-1)^2, heads] = relative_position_bias_table.  Pad to a multiple of
 * `window`, cyclic shift by `shift` (0 or window/2), window partition, softmax(q k^T / sqrt(32) +
 * relative position bias + shift mask (-100 across img_mask regions)) v, window reverse, shift
 * back and crop are folded into the addressing.  out fp32 [B,H,W,C] and / or (out_hi, out_lo)
 * bf16 [B,H,W,C]: the split operand planes of the result for the proj pvsg_linear_tc (hi = rn(v),
 * lo = rn(v - hi)).  C == heads*32, window <= 12, else PVSG_ERR_UNSUPPORTED. */
int pvsg_window_attention(const float* qkv, const float* qkv_bias, const float* bias_table, float* out,
                          void* out_hi, void* out_lo, int B, int H, int W, int C, int heads, int window,
                          int shift, void* stream);

/* mmdet/models/utils/transformer.py PatchMerging.forward :300-352 up to (not including) the
 * reduction linear: nn.Unfold(kernel 2, stride 2) channel order (c*4 + kh*2 + kw), zero padding
 * at the bottom / right for odd H / W ("corner" adaptive padding), LayerNorm over 4C.
 * x fp32 [B,H,W,C] -> y fp32 [B,ceil(H/2),ceil(W/2),4C]; gamma / beta [4C].  C % 4 == 0, C <= 1024. */
int pvsg_patch_merge_ln(const float* x, const float* gamma, const float* beta, float* y, int B, int H,
                        int W, int C, float eps, void* stream);

/* Joint histogram of a ground-truth instance-id map and a predicted panoptic map, per frame, for
 * the relation-set builder (reference: match_and_process_gt_tubes / calculate_iou,
 * utils/relation_matching.py:156-165,205-260 -- per (frame, GT object, same-class tube) it decodes
 * the tube's RLE mask and runs two full-frame logical passes).  gt int32 [B,H,W] holds object ids
 * (ids outside [0,G) are counted in row G); pan int32 [B,H,W] and seg_info [B,1+4Q] as written by
 * pvsg_panoptic_fuse (slots numbered as in pvsg_rle_events; column Q = pixels of no kept segment).
 * counts int32 [B,G+1,Q+1] (zeroed by the call): counts[b][g][s] = |gt g & slot s|; row sums =
 * GT areas, column sums = segment areas, so IoU(g,s) = c / (row_g + col_s - c) for every pair.
 * (G+1)*(Q+1) <= 51200 (shared-memory histogram). */
int pvsg_tube_overlap(const int32_t* gt, const int32_t* pan, const int32_t* seg_info, int B, int Q,
                      int H, int W, int G, int32_t* counts, void* stream);

/* ------------------------------------------------------------ relation head ----- */

/* y[n,c] = max_t x[n,t,c]  (base.py:50-51). */
int pvsg_max_over_time(const float* x, float* y, int N, int T, int C, void* stream);

/* Temporal taps of x [P,T,C] (zero padded cross-correlation along T, odd K, C % 4 == 0), the two baseline relation
 * models of models/relation_head/convolution.py:
 *   pvsg_temporal_fir    y[p,t,c] = sum_k w[k] x[p,t+k-K/2,c]      HandcraftedFilter's depthwise F.conv1d (:26-30)
 *   pvsg_temporal_unfold y[p,t,k*C+c] = x[p,t+k-K/2,c]  [P,T,K*C]  operand of Learnable1DConv's nn.Conv1d (:49-56),
 *                                                                   which then runs as ONE pvsg_linear[_tc] GEMM. */
int pvsg_temporal_fir(const float* x, const float* w, float* y, int P, int T, int C, int K, void* stream);
int pvsg_temporal_unfold(const float* x, float* y, int P, int T, int C, int K, void* stream);

/* PairProposalNetwork.forward (base.py:49-62), factorised: U = sub_tok W1[:, :F]^T + b1 and
 * V = obj_tok W1[:, F:]^T are precomputed [N,Hd] (two pvsg_linear calls);
 * pair[i,j] = b2 + sum_h w2[h] * relu(U[i,h] + V[j,h]) for i != j, 0 on the diagonal. */
int pvsg_pair_proposal(const float* U, const float* V, const float* w2, const float* b2,
                       float* pair, int N, int Hd, void* stream);

/* pick_top_pairs_eval (test_utils.py:4-22): diagonal -> -inf, the min(N*N, k) largest
 * entries in descending order (ties: lower flat index first); pairs int32 [k,2] = (s,o),
 * diagonal hits removed; n_out int32 [1]. k <= 1024 (radix select + ranking in one CTA). */
int pvsg_top_pairs(const float* pair, int N, int k, int32_t* pairs, int32_t* n_out, void* stream);

/* concatenate_sub_obj (train_utils.py:67-81) + PositionalEncoding add (transformer.py:77-81):
 * out[p,t,:] = [sub[s_p,t,:], obj[o_p,t,:]] + pe[t,:] (pe optional, [T,2F]). */
int pvsg_gather_pairs(const float* sub, const float* obj, const int32_t* pairs, const float* pe,
                      float* out, int P, int T, int F, void* stream);

/* Attention on the 5th-generation tensor cores: same contract as pvsg_attention_tc (K / V as split-bf16 operand planes,
 * head dim 32 or 128, byte mask + open-row rule, key splits merged through the workspace), but S = Q K^T and O = P V are
 * tcgen05.mma instructions with TMEM accumulators, K / V tiles arrive by TMA and V is consumed as an MN-major operand
 * (csrc/attention_t5.cu).  The K and V planes must share their strides (slices of one projection).  Replaces
 * models/relation_head/base.py:26-40 and transformer.py:35-56 (nn.TransformerEncoderLayer self-attention) and the
 * decoder cross-attention of models/mask2former/mask2former_head.py:457-468. */
int64_t pvsg_attention_t5_workspace_bytes(int B, int H, int Lq, int Lk, int D);
/* same, also returning lse [B,H,Lq] = log sum_k exp(scale q.k + mask) per row: what pvsg_attention_train_backward needs, so the
 * training forward of the decoder's attention runs on the tcgen05 kernel too. */
int pvsg_attention_t5_lse(const float* Q, const void* K_hi, const void* K_lo, const void* V_hi, const void* V_lo,
                          const uint8_t* mask, const int32_t* row_open, float* out, float* lse, void* ws, int B, int H, int Lq,
                          int Lk, int Dh, int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts,
                          int64_t o_bs, int64_t o_ts, float scale, void* stream);
int pvsg_attention_t5(const float* Q, const void* K_hi, const void* K_lo, const void* V_hi, const void* V_lo,
                      const uint8_t* mask, const int32_t* row_open, float* out, void* workspace, int B, int H, int Lq,
                      int Lk, int D, int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts,
                      int64_t o_bs, int64_t o_ts, float scale, void* stream);

/* ------------------------------------------------------------ IPS tracker path ----- */

/* F.interpolate(x, scale_factor=s, mode='bilinear') as models/unitrack/mask.py:36 calls it: the source coordinate of an
 * output pixel is (dst + 0.5) * scale - 0.5 with the CALLER's scale = 1 / s (ATen uses the given scale factor, not
 * in / out, when recompute_scale_factor is unset), output size floor(in * s).  Token-major [B,H,W,C], C % 4 == 0. */
int pvsg_bilinear_resize_scaled(const float* src, float* dst, int B, int IH, int IW, int OH, int OW, int C,
                                float scale_h, float scale_w, void* stream);

/* UniTrack reconstruction-similarity distance (models/unitrack/core/association/matching.py:194-238
 * `reconsdot_distance`, called from multitracker.py:27 `class_aware_distance`).  trk [ntrk,nst,d], det [ndet,nsd,d]:
 * mask-pooled appearance embeddings, position-major, zero padded to the longest (`get_track_feat`, :170-191).
 * cost [ntrk,ndet] = 1 - (cos(recons_trk, trk) + cos(recons_det, det)) / 2 with softmax temperature tmp (100). */
int64_t pvsg_reconsdot_workspace_bytes(int ntrk, int nst, int ndet, int nsd, int d);
int pvsg_reconsdot(const float* trk, const float* det, float* cost, void* workspace, int ntrk, int nst, int ndet,
                   int nsd, int d, float tmp, void* stream);

/* lap.lapjv(cost, extend_cost=True, cost_limit=thresh) as called by matching.py:29-40 `linear_assignment` (lap is a
 * third-party dependency of the reference, not vendored): exact assignment on the (n+m)-square extension whose extra
 * entries cost cost_limit / 2; x[i] = column of row i or -1, y[j] = row of column j or -1; +inf entries are never
 * matched.  n + m <= 256. */
int pvsg_lap_assign(const float* cost, int n, int m, double cost_limit, int32_t* x, int32_t* y, void* stream);

/* MinVIS tube linking, all frame pairs at once (models/mask2former_vps/mask2former_min_vis.py:244-258
 * `match_from_embds`: cosine cost + scipy linear_sum_assignment, one frame after the other).  The cost matrix of step t
 * depends only on the RAW query embeddings of frames t-1 and t (re-ordering the target permutes rows), so all T-1
 * problems are independent: pvsg_cosine_chain_cost -> cost [T-1,Q,Q] (rows: queries of frame t, columns: frame t+1),
 * pvsg_lap_square_batched -> x[t,i] = column assigned to row i (one warp per problem, exact, n <= 256); the clip's
 * permutations are then the running composition of the x[t] (host, T small integer gathers). */
int pvsg_cosine_chain_cost(const float* embeds, float* cost, int T, int Q, int C, void* stream);
int pvsg_lap_square_batched(const float* cost, int batch, int n, int32_t* x, int32_t* y, void* stream);
/* perms[0] = identity, perms[t][i] = x[t-1][perms[t-1][i]]: position i of the clip-long ordering holds query perms[t][i] of frame t. */
int pvsg_perm_chain(const int32_t* sigma, int32_t* perms, int T, int Q, void* stream);

/* ------------------------------------------------------------ training slice ----- */

/* mmcv.ops.point_sample as used by loss_single / _get_target_single (mask2former_video_head.py:175-178,262-267):
 * out[n,k] = bilinear sample (align_corners=False, zero padding) of maps[n] [H,W] at points [n,K,2] (x, y in [0,1];
 * points_per_map = 0: ONE [K,2] set shared by all maps).  _backward: grad_maps[n] = scatter of grad_out (zeroed first). */
int pvsg_point_sample(const float* maps, const float* points, float* out, int n, int H, int W, int K, int points_per_map,
                      void* stream);
int pvsg_point_sample_backward(const float* grad_out, const float* points, float* grad_maps, int n, int H, int W, int K,
                               int points_per_map, void* stream);

/* loss_mask + loss_dice of loss_single (:270-289) on sampled logits / targets [n,K]: sums[0] = sum of
 * binary_cross_entropy_with_logits terms (mmdet CrossEntropyLoss(use_sigmoid=True)), sums[1] = sum over rows of the
 * naive dice loss 1 - (2 sum(s t) + eps) / (sum s + sum t + eps) (mmdet DiceLoss(naive_dice=True)); grad (optional)
 * = bce_grad_scale * dBCE/dlogit + dice_grad_scale * ddice/dlogit. */
int pvsg_mask_point_losses(const float* logits, const float* targets, int n, int K, float dice_eps, float bce_grad_scale,
                           float dice_grad_scale, float* sums, float* grad, void* stream);

/* loss_cls of loss_single (:236-244): mmdet CrossEntropyLoss with class_weight [C] and per-row label_weight (may be
 * NULL): sums[0] = sum_r w[y_r] lw_r (logsumexp(x_r) - x_r[y_r]), sums[1] = sum_r w[y_r] (the avg_factor);
 * grad (optional) = grad_scale * w lw (softmax(x) - onehot). */
int pvsg_weighted_ce(const float* logits, const int64_t* labels, const float* class_weight, const float* label_weight,
                     int rows, int C, float grad_scale, float* sums, float* grad, void* stream);

/* mmdet MaskHungarianAssigner cost matrix (called from _get_target_single :181): cost[q,g] = -w_cls softmax(cls[q])[label_g]
 * + w_mask mean_k BCE cost (CrossEntropyLossCost, use_sigmoid) + w_dice DiceCost(pred_act, naive_dice, eps) over the
 * K sampled points; the assignment itself is scipy's linear_sum_assignment on the host, as in mmdet. */
int pvsg_mask_match_cost(const float* cls_logits, const int64_t* gt_labels, const float* pred_points, const float* gt_points,
                         int Q, int G, int C, int K, float w_cls, float w_mask, float w_dice, float dice_eps, float* cost,
                         void* stream);

/* mmcv MultiScaleDeformableAttnFunction.backward (the pixel decoder's attention, cfg :38-47): gradients of
 * pvsg_msda_forward w.r.t. value [B,N,H,32] (atomics, zeroed first), sampling_locations [B,Nq,H,L,P,2] and
 * attention_weights [B,Nq,H,L,P]. */
int pvsg_msda_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                       const float* sampling_locations, const float* attention_weights, const float* grad_out,
                       float* grad_value, float* grad_loc, float* grad_attn, int B, int64_t N, int64_t Nq, int H, int D,
                       int L, int P, void* stream);

/* ---- second training slice: backward of the decoder head (forward_train, mask2former_video_head.py:464-522).  Every
 * dense product of the backward pass is pvsg_linear again (dX = dY W, dW^T = X^T dY on transposed copies). ---- */

/* nn.LayerNorm backward over the last axis of x [M,C]: dx [M,C]; dgamma / dbeta [C] = sums over the rows (zeroed here). */
int pvsg_layernorm_backward(const float* x, const float* gamma, const float* dy, float* dx, float* dgamma, float* dbeta,
                            int64_t M, int C, float eps, void* stream);
/* dx = dy where the forward OUTPUT y of a ReLU-fused layer was positive, else 0. */
int pvsg_relu_backward(const float* dy, const float* y, float* dx, int64_t n, void* stream);
/* out[n] = sum_m x[m*ld + n]: bias gradients, d level_embed, gradient of a batch broadcast (out zeroed here). */
int pvsg_colsum(const float* x, float* out, int64_t M, int N, int64_t ld, void* stream);
/* nn.MultiheadAttention core as the decoder calls it (mask2former_head.py:457-468) in training: out = softmax(scale q k^T
 * + mask) v per head, exact fp32, also returning the row log-sum-exp lse [B,H,Lq] for the backward.  q [B,Lq,H*D],
 * k / v [B,Lk,H*D] with batch / row strides in floats (multiples of 4), D = 32; mask uint8 [B,Lq,Lk] (non-zero = blocked,
 * shared by the heads) with row_open [B,Lq] = open keys per row (a row with none ignores the mask, :451-452); both may be
 * NULL.  _backward: dq [B,Lq,H*D], dk / dv [B,Lk,H*D] contiguous; `dout` laid out like `out`; delta [B,H,Lq] scratch. */
int pvsg_attention_train_forward(const float* q, const float* k, const float* v, const uint8_t* mask, const int32_t* row_open,
                                 float* out, float* lse, int B, int H, int Lq, int Lk, int D, int64_t q_bs, int64_t q_rs,
                                 int64_t k_bs, int64_t k_rs, int64_t v_bs, int64_t v_rs, int64_t o_bs, int64_t o_rs, float scale,
                                 void* stream);
int pvsg_attention_train_backward(const float* q, const float* k, const float* v, const uint8_t* mask, const int32_t* row_open,
                                  const float* out, const float* dout, const float* lse, float* delta, float* dq, float* dk,
                                  float* dv, int B, int H, int Lq, int Lk, int D, int64_t q_bs, int64_t q_rs, int64_t k_bs,
                                  int64_t k_rs, int64_t v_bs, int64_t v_rs, int64_t o_bs, int64_t o_rs, float scale, void* stream);

/* ---- third training slice: backward of the pixel decoder (mmdet MSDeformAttnPixelDecoder, cfg :27-59).  Convolutions
 * go backward through the forward engine: dX of a 3x3 conv = pvsg_conv2d_nhwc of dY with the flipped, transposed filter;
 * dW = one pvsg_linear per filter tap on shifted copies of the input. ---- */

/* nn.GroupNorm backward on token-major x [B,HW,C] (mmcv ConvModule norm, optionally followed by ReLU: relu != 0 masks dy
 * where the forward output was <= 0): dx, dgamma / dbeta [C] (zeroed here); stats: scratch [B*G*4] floats. */
int pvsg_groupnorm_nhwc_backward(const float* x, const float* gamma, const float* beta, const float* dy, float* dx, float* dgamma,
                                 float* dbeta, float* stats, int B, int64_t HW, int C, int G, float eps, int relu, void* stream);
/* adjoint of pvsg_bilinear_resize_nhwc (F.interpolate bilinear, align_corners=False, the FPN top-down step
 * msdeformattn_pixel_decoder.py L0): dsrc [B,IH,IW,C] (zeroed here) += scatter of dout [B,OH,OW,C]. */
int pvsg_bilinear_resize_nhwc_backward(const float* dout, float* dsrc, int B, int IH, int IW, int OH, int OW, int C, void* stream);
/* Operand preparation of the weight-gradient GEMMs (dW = dZ^T X, reduction over the tokens): rows r = (b, oh, ow) of a
 * token-major view (element strides sb, sh, sw; channels contiguous; optional `add` with the same layout) -> bf16 (hi, lo)
 * planes [C, ldt], column r.  The planes feed pvsg_linear_tc_batched as K-chunk views (split-K), summed by pvsg_colsum. */
int pvsg_transpose_split(const float* x, const float* add, void* hi, void* lo, int64_t T, int C, int OH, int OW, int64_t sb,
                         int64_t sh, int64_t sw, int64_t ldt, void* stream);
/* backward of pvsg_maxpool3x3s2_nhwc (ResNet stem max pooling): dx [B,H,W,C] (zeroed here); the gradient of a window goes to
 * its first maximum in scan order (ATen max_pool2d_with_indices_backward). */
int pvsg_maxpool3x3s2_nhwc_backward(const float* x, const float* dy, float* dx, int B, int H, int W, int C, void* stream);
/* proj [B,Nq,H*L*P*3] (see pvsg_msda_fused_forward) -> sampling locations [B,Nq,H,L,P,2] and attention weights [B,Nq,H,L,P],
 * the explicit tensors pvsg_msda_backward wants; _backward maps their gradients back to dproj. */
int pvsg_msda_proj_expand(const float* proj, const float* ref, const int64_t* spatial_shapes, float* loc, float* aw, int B,
                          int64_t Nq, int H, int L, int P, void* stream);
int pvsg_msda_proj_backward(const float* aw, const float* dloc, const float* daw, const int64_t* spatial_shapes, float* dproj,
                            int B, int64_t Nq, int H, int L, int P, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVSG_H_ */
