/*
 * fcp_b200.h — C ABI of libfcpb200.so, the B200 (sm_100a) implementation of the face-crop-plus hot path
 * (detect -> align -> [enhance] -> parse).
 *
 * The reference (mantasu/face-crop-plus, pure Python) has no FFI of its own: the path sits behind Python
 * objects.  Each entry point below replaces one of those Python call sites; the host mirror in
 * face_crop_plus_b200/ binds them with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - every function returns an int status (FCP_OK == 0); fcp_last_error(ctx) gives the message;
 *  - no exceptions, no ownership transfer: inputs are borrowed, outputs are caller-allocated with explicit capacities;
 *  - every data pointer may be HOST or DEVICE memory (queried with cudaPointerGetAttributes); host buffers are
 *    staged through pinned memory inside the call, so host<->device copies are part of the call;
 *  - images / crops are uint8 NHWC RGB; landmarks are float32 (x, y) pairs; matrices are float64 2x3 row-major;
 *  - a context is bound to one CUDA device and one stream and is thread-compatible (one context per worker thread).
 */
#ifndef FCP_B200_H
#define FCP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCP_OK 0
#define FCP_ERR_INVALID 1   /* bad argument / unknown key / unsupported shape                           */
#define FCP_ERR_CUDA 2      /* a CUDA runtime call or kernel failed                                      */
#define FCP_ERR_STATE 3     /* model not finalized, weights missing                                      */
#define FCP_ERR_CAPACITY 4  /* an output capacity was too small (counts are still written; call again)   */

typedef struct fcp_ctx fcp_ctx;

/* model ids: models/retinaface.py:10, models/bise.py:8, models/rrdb.py:8 */
enum { FCP_MODEL_RETINAFACE = 0, FCP_MODEL_BISENET = 1, FCP_MODEL_RRDBNET = 2 };
/* RetinaFace.take_by_strategy, models/retinaface.py:306-408 */
enum { FCP_STRATEGY_ALL = 0, FCP_STRATEGY_BEST = 1, FCP_STRATEGY_LARGEST = 2 };
/* cv2.BORDER_* values reachable through Cropper(padding=...), cropper.py:512 */
enum { FCP_BORDER_CONSTANT = 0, FCP_BORDER_REPLICATE = 1, FCP_BORDER_REFLECT = 2, FCP_BORDER_WRAP = 3,
       FCP_BORDER_REFLECT_101 = 4 };
/* activation codes of fcp_conv2d */
enum { FCP_ACT_NONE = 0, FCP_ACT_RELU = 1, FCP_ACT_LRELU = 2, FCP_ACT_SIGMOID = 3 };

/* ---- lifetime --------------------------------------------------------------------------------------- */
int fcp_create(int device, fcp_ctx** out);
void fcp_destroy(fcp_ctx* ctx);
const char* fcp_last_error(const fcp_ctx* ctx);
/* ABI/version string of the library ("fcp_b200 <semver> sm_100a") */
const char* fcp_version(void);
/* use an existing cudaStream_t (e.g. torch's current stream) instead of the context's own stream */
int fcp_set_stream(fcp_ctx* ctx, void* cuda_stream);
/* block until all work queued by this context has finished */
int fcp_sync(fcp_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py reports it as gpu_launches) */
int64_t fcp_launch_count(const fcp_ctx* ctx);
/* images per detector micro-batch / faces per parser micro-batch (bounds the activation arena; default 32 / 64) */
int fcp_set_micro_batch(fcp_ctx* ctx, int detect_images, int parse_faces);
/* convolution kernel used by the model graphs: 2 = tcgen05 3xFP16 block-scaled split (default), 1 = tcgen05 3xTF32 split,
 * 0 = CUDA-core fp32; all three are fp32-accurate (error vs fp64 <= that of an fp32 FMA chain; tests/test_gpu_parity.py).
 * 3 = opt-in fast mode: the block-scaled FP16 kernel without its two correction terms (one tensor-core pass, 11 significant
 * bits per operand, the accuracy class of cuDNN's TF32 path) - NOT inside the parity bar, never the default.
 * The environment variable FCP_CONV_IMPL sets the default of new contexts. */
int fcp_set_conv_impl(fcp_ctx* ctx, int impl);

/* per-kernel timing for bench.py's roofline line: when enabled, every convolution launch is bracketed by CUDA events
 * on the context stream.  fcp_profile_read synchronizes and returns, accumulated since the last read:
 * out[0] = convolution kernel time (ms), out[1] = convolution launches, out[2] = algorithmic convolution FLOPs
 * (2*M*Cout*K of the true, un-padded shapes), out[3] = algorithmic bytes (activations in+out, weights once per launch) */
int fcp_profile(fcp_ctx* ctx, int enable);
int fcp_profile_read(fcp_ctx* ctx, double* out4);
/* while profiling is enabled every stage of the path is bracketed by an event pair as well; returns (and resets) the
 * accumulated milliseconds: out[0] detector network, [1] decode + NMS + strategy, [2] enhance, [3] un-pad + solve + warp,
 * [4] parser network, [5] parser tail (resize + argmax + histogram), [6] metadata all-gather (side stream), [7] ingest */
int fcp_profile_stages(fcp_ctx* ctx, double* out_ms8);

/* ---- weights: replaces LoadMixin.load / get_weights (models/_layers.py:16-35) ---------------------------
 * Feed every entry of the reference state_dict (same keys, float32, host memory, OIHW conv weights), then
 * finalize: BN running stats are folded into per-channel scale/shift, conv weights are re-packed for the
 * kernels and uploaded.  num_batches_tracked entries may be skipped.  rrdb_blocks: RRDB_trunk length (23). */
int fcp_load_tensor(fcp_ctx* ctx, int model, const char* key, const float* host_data, const int64_t* shape, int ndim);
int fcp_finalize(fcp_ctx* ctx, int model, int rrdb_blocks);

/* ---- ingest: replaces utils.as_batch (utils.py:273-342) ------------------------------------------------
 * image_ptrs[i] u8 [hs[i],ws[i],3] RGB (host or device).  Every image is resized to fit (size_w, size_h) keeping its
 * aspect ratio - OpenCV's INTER_AREA arithmetic when max(h,w) > max(size), its INTER_CUBIC arithmetic otherwise, a plain
 * copy when the size already fits (utils.py:320,334) - and centred with copyMakeBorder semantics (utils.py:335;
 * border_mode = FCP_BORDER_*).  out_batch u8 [n,size_h,size_w,3] (host or device); out_unscales f64 [n] and
 * out_paddings i32 [n,4] = (top,bottom,left,right) are HOST arrays (either may be NULL).
 * INTER_CUBIC exists in two arithmetics (fcp_set_cubic_mode): 1 (default) = floating point, what the opencv-python x86
 * wheels compute (they route 8-bit cubic resizes to Intel IPP; restated in float64: one grey level off in < 1e-5 of the
 * bytes); 0 = OpenCV's own 11-bit fixed-point code (bit-exact vs cv2 with cv2.ipp.setUseIPP(False)).  FCP_CUBIC=opencv
 * selects 0 for new contexts.  INTER_AREA and the borders are bit-exact vs cv2 either way. */
int fcp_set_cubic_mode(fcp_ctx* ctx, int floating_point);
int fcp_as_batch(fcp_ctx* ctx, const uint8_t* const* image_ptrs, const int32_t* hs, const int32_t* ws, int n,
                 int size_w, int size_h, int border_mode, uint8_t* out_batch, double* out_unscales,
                 int32_t* out_paddings);

/* ---- detect: replaces RetinaFace.predict (models/retinaface.py:410-470) ---------------------------------
 * images  u8 [n,h,w,3] RGB (what utils.as_batch produces, before as_tensor's float conversion; h,w % 32 == 0 not
 *         required).  Outputs, ordered by image then by the strategy's order, capacity max_faces:
 * out_landmarks f32 [max_faces,5,2] (batch pixel coords), out_indices i32 [max_faces] image index per face,
 * out_boxes f32 [max_faces,4] x1y1x2y2, out_scores f32 [max_faces], out_anchors i32 [max_faces] prior index,
 * out_count i32 [1] number of faces (written even on FCP_ERR_CAPACITY).  Any output pointer except out_count
 * and out_landmarks/out_indices may be NULL. */
int fcp_detect(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, float vis_threshold, float nms_threshold,
               int strategy, int max_faces, float* out_landmarks, int32_t* out_indices, float* out_boxes,
               float* out_scores, int32_t* out_anchors, int32_t* out_count);
/* raw head outputs of RetinaFace.forward before softmax (models/retinaface.py:137-142):
 * out_heads f32 [n, A, 16] = (cls[2], box[4], ldm[10]) per prior, priors in the reference order (level,row,col,anchor) */
int fcp_detect_heads(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, float* out_heads);
/* decode + threshold + NMS + strategy on given head outputs (models/retinaface.py:146-408, _layers.py:41-62) */
int fcp_detect_post(fcp_ctx* ctx, const float* heads, int n, int h, int w, float vis_threshold, float nms_threshold,
                    int strategy, int max_faces, float* out_landmarks, int32_t* out_indices, float* out_boxes,
                    float* out_scores, int32_t* out_anchors, int32_t* out_count);

/* ---- align: replaces Cropper.crop_align (cropper.py:441-552) --------------------------------------------
 * images u8 [n,h,w,3]; paddings i32 [n,4] = (top,bottom,left,right) or NULL; indices i32 [f]; landmarks f32 [f,5,2]
 * (already un-padded, cropper.py:822); target f32 [5,2] (cropper.py:392-439).  out_crops u8 [f,out_h,out_w,3]
 * is written for EVERY face slot (invalid faces are zero-filled; the host drops them like cropper.py:529-531),
 * out_matrices f64 [f,2,3] (may be NULL), out_valid u8 [f] (may be NULL). */
int fcp_align(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, const int32_t* paddings,
              const int32_t* indices, const float* landmarks, int f, const float* target, int out_w, int out_h,
              int border_mode, int allow_skew, uint8_t* out_crops, double* out_matrices, uint8_t* out_valid);
/* same, for a ragged list of images (landmarks-only path, cropper.py:796-813): image_ptrs[i] is u8 [hs[i],ws[i],3] */
int fcp_align_list(fcp_ctx* ctx, const uint8_t* const* image_ptrs, const int32_t* hs, const int32_t* ws, int n,
                   const int32_t* paddings, const int32_t* indices, const float* landmarks, int f,
                   const float* target, int out_w, int out_h, int border_mode, int allow_skew, uint8_t* out_crops,
                   double* out_matrices, uint8_t* out_valid);

/* N-point -> 5-point landmark reduction: replaces get_ldm_slices + the slice means of Cropper.process_batch
 * (utils.py:90-168, cropper.py:828-831).  landmarks f32 [f,k,2] with k in {5,12,17,21,29,49,68,98,106} (any other k:
 * FCP_ERR_INVALID, where the reference raises ValueError); out f32 [f,5,2] = (left eye, right eye, nose, mouth corners). */
int fcp_reduce_landmarks(fcp_ctx* ctx, const float* landmarks, int f, int k, float* out);

/* ---- parse: replaces BiSeNet.predict (models/bise.py:327-418) -------------------------------------------
 * crops u8 [f,h,w,3]; out_labels u8 [f,h,w] (argmax class 0..18), out_hist i32 [f,19] pixel count per class
 * (either may be NULL).  Grouping by thresholds (bise.py:214-325) is integer work on out_hist done by the host;
 * fcp_masks writes the 0/255 masks of one mask group: out_masks u8 [f,h,w] = 255 where class_lut[label] != 0. */
int fcp_parse(fcp_ctx* ctx, const uint8_t* crops, int f, int h, int w, uint8_t* out_labels, int32_t* out_hist);
/* BiSeNet logits at 1/8 resolution before the final upsample: out_logits f32 [f,19,64,64] NCHW (bise.py:211) */
int fcp_parse_logits(fcp_ctx* ctx, const uint8_t* crops, int f, int h, int w, float* out_logits);
/* the tail only: bilinear(align_corners) to 512x512 -> nearest to (h,w) -> argmax, on given logits (bise.py:212,394) */
int fcp_parse_tail(fcp_ctx* ctx, const float* logits, int f, int h, int w, uint8_t* out_labels, int32_t* out_hist);
int fcp_masks(fcp_ctx* ctx, const uint8_t* labels, int f, int h, int w, const uint8_t* class_lut19, uint8_t* out_masks);

/* grouping: replaces BiSeNet.group_by_attributes / group_by_masks (bise.py:214-325) - integer work on the histogram and
 * the labels, no host round trip.  hist i32 [f,19], labels u8 [f,h,w] (only read when out_masks != NULL).
 * Attribute group g = attr_codes[attr_offsets[g] .. attr_offsets[g+1]) (signed class ids as in Cropper(attr_groups=...):
 * a > 0: more than attr_threshold pixels of class a; otherwise at most attr_threshold pixels of class |a|), joined by AND
 * when join_and != 0 (attr_join_by_and, bise.py:184), else OR.  Mask group m = the classes c with mask_lut[m*19+c] != 0;
 * a face belongs to it when more than mask_threshold pixels fall into those classes (bise.py:316).  n_mask <= 32.
 * out_attr u8 [n_attr,f], out_mask u8 [n_mask,f] membership flags; out_masks u8 [n_mask,f,h,w] = 0/255 for EVERY face
 * (NULL = skip; the caller keeps the members, like mask[inds] at bise.py:317). */
int fcp_group(fcp_ctx* ctx, const uint8_t* labels, const int32_t* hist, int f, int h, int w, const int32_t* attr_codes,
              const int32_t* attr_offsets, int n_attr, int attr_threshold, int join_and, const uint8_t* mask_lut,
              int n_mask, int mask_threshold, uint8_t* out_attr, uint8_t* out_mask, uint8_t* out_masks);

/* ---- enhance: replaces RRDBNet.predict / forward (models/rrdb.py:64-146) --------------------------------
 * images f32 [n,3,h,w] NCHW, 0..255 (the tensor Cropper hands over, cropper.py:835); enhanced IN PLACE where
 * do_enhance[i] != 0 (gate computed by the host from landmarks, rrdb.py:125-140; NULL = all). */
int fcp_enhance(fcp_ctx* ctx, float* images, int n, int h, int w, const uint8_t* do_enhance);
/* RRDBNet.forward: x f32 [n,3,h,w] in [0,1] -> out f32 [n,3,4h,4w] */
int fcp_enhance_forward(fcp_ctx* ctx, const float* x, int n, int h, int w, float* out);
/* the same predict() on a uint8 NHWC batch [n,h,w,3] (what Cropper holds before as_tensor, cropper.py:817): the
 * reference's result round(clamp(bicubic(x4, 1/4), 0, 1) * 255) is integral, so uint8 in/out loses nothing and the
 * float32 NCHW copies (12.6 MB per 1024x1024 image, each way) never exist.  In place where do_enhance[i] != 0. */
int fcp_enhance_u8(fcp_ctx* ctx, uint8_t* images, int n, int h, int w, const uint8_t* do_enhance);
/* the gate of RRDBNet.predict (rrdb.py:124-141) on the device: out_gate[i] = 1 iff image i has faces and the float32 mean
 * of (x4-x0)*(y4-y0) / (h*w) over them is <= min_face_factor.  landmarks f32 [f,5,2], indices i32 [f] ascending. */
int fcp_enhance_gate(fcp_ctx* ctx, const float* landmarks, const int32_t* indices, int f, int n, int h, int w,
                     float min_face_factor, uint8_t* out_gate);
/* enable != 0 makes fcp_pipeline run the enhancement stage between detection and alignment (cropper.py:833-836: gate on
 * the device from the un-padded landmarks, RRDBNet on the gated images of the uint8 batch, crops warped from the enhanced
 * images); off by default.  Needs the RRDBNet weights. */
int fcp_set_enhance(fcp_ctx* ctx, int enable, float min_face_factor);

/* ---- whole path: the detect branch of Cropper.process_batch (cropper.py:815-847) in one call ------------
 * detect -> un-pad -> align -> parse with no host round trip between the stages.  Capacities as in fcp_detect;
 * out_crops u8 [max_faces,out_h,out_w,3], out_labels u8 [max_faces,out_h,out_w] (NULL = skip parsing),
 * out_hist i32 [max_faces,19], out_matrices f64 [max_faces,2,3], out_valid u8 [max_faces]. */
int fcp_pipeline(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, const int32_t* paddings,
                 float vis_threshold, float nms_threshold, int strategy, const float* target, int out_w, int out_h,
                 int border_mode, int allow_skew, int max_faces, float* out_landmarks, int32_t* out_indices,
                 int32_t* out_count, uint8_t* out_crops, double* out_matrices, uint8_t* out_valid,
                 uint8_t* out_labels, int32_t* out_hist);

/* ---- multi-GPU: batch sharding with ONE collective -----------------------------------------
 * One process and one context per GPU; every rank runs the path on its shard of the images.  The only data that crosses
 * devices is an all-gather of fixed-size per-face records over NCCL: 20 float64 per face = landmarks[10], GLOBAL image
 * index, matrix[6], valid, 2 reserved; a rank's block is [cap + 1][20] with its face count in [cap][0].
 * The reference has no multi-device code (one torch.device per Cropper, cropper.py:336-337).
 * fcp_comm_unique_id: rank 0 creates the 128-byte NCCL id, the host side broadcasts it (any channel);
 * fcp_comm_init: collective over all ranks.  NCCL is bound at run time (libnccl.so.2 of the process, or FCP_NCCL_LIB). */
int fcp_comm_unique_id(fcp_ctx* ctx, void* out_id128);
int fcp_comm_init(fcp_ctx* ctx, int rank, int world, const void* id128);
void fcp_comm_destroy(fcp_ctx* ctx);
/* explicit call: packs `count` local faces on the device and all-gathers; out_records f64 [world][cap + 1][20] */
int fcp_allgather_meta(fcp_ctx* ctx, const float* landmarks, const int32_t* indices, const double* matrices,
                       const uint8_t* valid, int count, int cap, int index_base, double* out_records);
/* fused form: while out_records (DEVICE memory, [world][cap + 1][20]) is set, every fcp_pipeline call packs its face
 * records right after the align stage and all-gathers them on a side stream, overlapped with the parser; the result is
 * complete when fcp_pipeline returns.  index_base = global index of this rank's first image.  NULL turns it off. */
int fcp_set_gather(fcp_ctx* ctx, double* out_records, int cap, int index_base);

/* ---- kernel-level test hook: one fused convolution (what nn.Conv2d + BatchNorm2d + activation do in the
 * reference graphs).  x f32 NHWC [n,h,w,cin]; weight f32 OIHW host [cout,cin,k,k]; scale/shift f32 [cout] host
 * (NULL = 1/0); residual f32 NHWC [n,ho,wo,cout] added before the activation (NULL = none);
 * out f32 NHWC [n,ho,wo,cout].  impl: 0 = CUDA-core fp32 kernel, 1 = tcgen05 3xTF32 kernel, 2 = tcgen05 3xFP16 block-scaled kernel. */
int fcp_conv2d(fcp_ctx* ctx, const float* x, int n, int h, int w, int cin, const float* weight, int cout, int k,
               int stride, int pad, const float* scale, const float* shift, const float* residual, int act,
               float slope, int impl, float* out);

#ifdef __cplusplus
}
#endif
#endif /* FCP_B200_H */
