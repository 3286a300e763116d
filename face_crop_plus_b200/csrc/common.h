// Internal declarations shared by the runtime and the kernels of libfcpb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/fcp_b200.h"

namespace fcp {

// ----------------------------------------------------------------------------------------------------------
// activation tensor view: NHWC float32; `cs` is the channel stride of the underlying buffer (>= c) and `co`
// the channel offset of this view inside it, so concatenations are written in place (cat-free).
// ----------------------------------------------------------------------------------------------------------
struct Tensor {
    float* p = nullptr;
    int n = 0, h = 0, w = 0, c = 0;
    int cs = 0, co = 0;
    size_t pixels() const { return (size_t)n * h * w; }
    Tensor slice(int off, int count) const {
        Tensor t = *this;
        t.co = co + off;
        t.c = count;
        return t;
    }
};

// ----------------------------------------------------------------------------------------------------------
// one fused convolution: out = post(act(conv(in) * scale + shift + res1)) ...
// ----------------------------------------------------------------------------------------------------------
struct ConvWeights {
    int cout = 0, cin = 0, k = 1;
    int kh = 0, kw = 0;        // != 0: non-square kernel (tcgen05 kernel only); the stems run as a 7x1 conv over row patches
    int alg_k = 0;             // != 0: K of the reference convolution this launch implements (FLOP accounting of fcp_profile)
    int cout_pad = 0;          // columns of the packed matrix (multiple of 32)
    float* w_kn = nullptr;     // device [k*k*cin][cout_pad]  (row = (r*k+s)*cin + c), CUDA-core kernel
    float* w_hi = nullptr;     // device [cout_pad][k*k*cin] K-major, tf32-truncated part   (tcgen05 kernel)
    float* w_lo = nullptr;     // device [cout_pad][k*k*cin] K-major, residual part          (tcgen05 kernel)
    // tcgen05 kernel, f16x3 mode: w * scale * 2^w_exp split into fp16 hi + lo, K-major [cout_pad][taps * cin_p rounded up to 64]
    void* h_hi = nullptr;
    void* h_lo = nullptr;
    int cin_p = 0;             // cin rounded up to 32 (K order: tap, channel; one K-block = two 32-channel units)
    int w_exp = 0;
    float* scale = nullptr;    // device [cout_pad]
    float* shift = nullptr;    // device [cout_pad]
};

struct ConvOp {
    Tensor in, out;
    const ConvWeights* wt = nullptr;
    int stride = 1, pad = 0;
    int stride_w = 0, pad_w = -1;   // horizontal stride / padding when they differ from the vertical ones (0 / -1: same)
    Tensor in2; int in2_stride = 1;   // 1x1 stride-1 convs on the tensor-core kernel: second K source - channels [in.c, in.c + in2.c) of the
                               // weight read `in2` at pixel (ho * in2_stride, wo * in2_stride) (a bottleneck's shortcut conv folded into conv3)
    int up_in = 0;             // read the input through a nearest x2 upsample (logical size = 2x physical)
    int act = FCP_ACT_NONE;
    float slope = 0.f;
    int out_rs = 0; size_t out_is = 0;   // != 0: row / image stride of `out` in floats when they are not w*cs / h*w*cs (a strided output view:
                               // RRDBNet's upsampling convs write one output-pixel parity class each; tensor-core kernel, TMA epilogue only)
    int out_add = 0;           // out += result instead of out = result (tensor-core kernel with the TMA epilogue only; RRDB skip, graphs.cu)
    int act_cols = 1 << 30;    // the activation applies to output channels < act_cols only (tcgen05 kernel; RRDBNet source-major passes)
    const float* res1 = nullptr; int res1_cs = 0, res1_co = 0;              // added before the activation
    float post_scale = 1.f;
    const float* res2 = nullptr; int res2_cs = 0, res2_co = 0;              // added after act*post_scale
    int res2_h = 0, res2_w = 0;                                             // != 0: res2 is [n,res2_h,res2_w,*], read through a nearest resize
    float post_scale2 = 1.f;
    const float* res3 = nullptr; int res3_cs = 0, res3_co = 0;              // added after (..)*post_scale2
    int a_exact = 0;           // the input values are small integers (u8 - mean): the split's low part is zero
    // detector stem straight from the uint8 RGB batch [n, stem_h, stem_w, 3] (tcgen05 f16x3 kernel only; `in` is unused)
    const uint8_t* stem_src = nullptr; int stem_h = 0, stem_w = 0;
    int impl = 0;              // 0 CUDA-core fp32, 1 tcgen05 3xTF32, 2 tcgen05 3xFP16 block-scaled, 3 = 2 without the correction terms
};

// ----------------------------------------------------------------------------------------------------------
// simple stream-ordered arena: first-fit free list over one big cudaMalloc
// ----------------------------------------------------------------------------------------------------------
class Arena {
public:
    ~Arena();
    bool reserve(size_t bytes);           // (re)allocates the slab if it is smaller than `bytes`
    void* alloc(size_t bytes);            // returns nullptr when exhausted
    void free(void* p);
    void reset();
    void set_plan_mode(bool on);          // plan mode: unlimited virtual space, fake addresses, records high_water
    bool plan_mode() const { return plan_; }
    size_t capacity() const { return cap_; }
    size_t high_water() const { return high_; }
private:
    struct Block { size_t off, size; bool used; };
    char* base_ = nullptr;
    size_t cap_ = 0, high_ = 0;
    bool plan_ = false;
    std::vector<Block> blocks_;
};

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

struct Model {
    std::map<std::string, HostTensor> host;          // as fed by fcp_load_tensor
    std::map<std::string, ConvWeights> conv;         // finalized (folded + packed) convolutions by name
    std::map<std::string, std::vector<float>> vec;   // small host-side vectors (e.g. folded BN for 1x1 GEMVs)
    bool finalized = false;
    int rrdb_blocks = 23;
};

}  // namespace fcp

struct fcp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string error;
    int64_t launches = 0;
    int det_mb = 32, par_mb = 64;
    fcp::Model models[3];
    fcp::Arena arena;          // activations
    fcp::Arena scratch;        // staging of host inputs/outputs, candidate buffers
    std::vector<void*> device_allocs;   // weights etc. freed at destroy
    void* pinned = nullptr; size_t pinned_bytes = 0;
    int sm_count = 148;
    int use_tc = 2;            // conv implementation of the model graphs: 2 tcgen05 3xFP16 block-scaled (default),
                               // 1 tcgen05 3xTF32, 0 CUDA-core fp32
    // profiling (fcp_profile): event pairs around conv launches + algorithmic work counters
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;   // start0, stop0, start1, stop1, ...
    size_t prof_used = 0;
    double prof_flops = 0, prof_bytes = 0;
    struct ProfRec { int m, cout, cin, k, stride, tc; };
    std::vector<ProfRec> prof_recs;     // one per event pair (FCP_TRACE=1 prints a per-shape table)
    // per-stage event pairs (fcp_profile_stages): detect net, detect post, enhance, align, parse net, parse tail, gather, ingest
    struct StageRec { int stage; cudaEvent_t a, b; };
    std::vector<StageRec> stage_recs;
    size_t stage_used = 0;
    // INTER_CUBIC arithmetic of fcp_as_batch: true = floating point (what IPP-enabled cv2 builds compute), false = OpenCV's
    // own 11-bit fixed-point code
    bool cubic_float = true;
    // RRDBNet stage of fcp_pipeline (fcp_set_enhance): on/off + the min_face_factor threshold of rrdb.py:141 (any float: the
    // face factor of a mirrored / degenerate landmark set is negative)
    bool enh_enabled = false;
    float enh_threshold = 0.001f;
    // multi-GPU metadata all-gather (comm.cu): NCCL communicator + side stream; gather_out != nullptr makes fcp_pipeline
    // pack and all-gather its face records on comm_stream while the parser runs
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t comm_ready = nullptr, comm_done = nullptr;
    double* gather_out = nullptr; int gather_cap = 0, gather_base = 0;
    double* gather_send = nullptr; size_t gather_send_cap = 0;
    // fcp_pipeline with HOST images: the batch is copied H2D per detector micro-batch on a second stream, one micro-batch
    // ahead of the compute stream (the hook runs at the top of every detector micro-batch)
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    std::function<int(int)> on_microbatch;   // called with the micro-batch number at the top of every detector micro-batch
    std::vector<int> mb_sched;               // detector micro-batch sizes installed by fcp_pipeline for that call (empty: uniform)
};

namespace fcp {

int fail(fcp_ctx* ctx, int code, const std::string& msg);

enum Stage { ST_DETECT_NET = 0, ST_DETECT_POST, ST_ENHANCE, ST_ALIGN, ST_PARSE_NET, ST_PARSE_TAIL, ST_GATHER, ST_INGEST, ST_COUNT };
// event pair around a stage while profiling is on (no-op otherwise); the pair is recorded on `stream` (default: ctx->stream)
struct StageScope {
    StageScope(fcp_ctx* ctx, int stage, cudaStream_t stream = nullptr, bool use_stream = false);
    ~StageScope();
    fcp_ctx* ctx; cudaStream_t stream; cudaEvent_t end = nullptr;
};
// packs this rank's face records and all-gathers them on the communicator's side stream (comm.cu); no-op without a comm
int gather_meta_async(fcp_ctx* ctx, const float* landmarks, const int32_t* face_img, const double* matrices, const uint8_t* valid,
                      const int32_t* face_count, int cap, int index_base, double* out_records);
int gather_meta_join(fcp_ctx* ctx);

#define FCP_CUDA(ctx, expr)                                                                            \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return fcp::fail(ctx, FCP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    } while (0)

#define FCP_TRY(expr)                  \
    do {                               \
        int _s = (expr);               \
        if (_s != FCP_OK) return _s;   \
    } while (0)

#define FCP_KERNEL_CHECK(ctx)                                                                          \
    do {                                                                                               \
        (ctx)->launches++;                                                                             \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess)                                                                         \
            return fcp::fail(ctx, FCP_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    } while (0)

// ---- kernels (one launcher per .cu) -----------------------------------------------------------------------
int launch_conv_ffma(fcp_ctx* ctx, const ConvOp& op);
int launch_conv_tc(fcp_ctx* ctx, const ConvOp& op);
bool conv_tc_supported(const ConvOp& op);
bool conv_tc_stem_supported(const void* images, int h, int w);
int run_conv(fcp_ctx* ctx, const ConvOp& op);

// stem: 7x7/s2 conv (Cin=3 -> 64) + scale/shift + ReLU.  mode 0: src = u8 RGB NHWC, flipped to BGR and mean-subtracted
// (retinaface.py:450-451); mode 1: src = f32 NHWC 3-channel
int launch_stem7(fcp_ctx* ctx, const void* src, int mode, int n, int h, int w, const float* w_kn /*[147][64]*/,
                 const float* scale, const float* shift, Tensor out);
// 5-point means of an N-point annotation; bounds = {lo0, hi0, ..., lo4, hi4} (utils.py:90-132)
int launch_reduce_landmarks(fcp_ctx* ctx, const float* lms, int f, int k, const int bounds[10], float* out);
// batch ingest (utils.as_batch): resize (OpenCV INTER_AREA / INTER_CUBIC arithmetic) + centred padding of a ragged image list;
// dev_ptrs[i] = device u8 [hs[i], ws[i], 3]; out = device u8 [n, size_h, size_w, 3]; unscales / paddings are HOST outputs
int launch_ingest(fcp_ctx* ctx, const uint8_t* const* dev_ptrs, const int32_t* hs, const int32_t* ws, int n, int size_w,
                  int size_h, int border_mode, uint8_t* out, double* unscales, int32_t* paddings);
// stem, tensor-core route: rows[n][h][wo][32] = for every INPUT row and every output column wo the 7 horizontal taps x 3
// channels (channel s*3 + c = tap s of conv channel c, 21 used, rest 0; zero outside the image) in fp32, same input
// conventions as launch_stem7.  The 7x7/s2 conv is then a 7x1 conv with stride (2,1) over `rows` (conv_tc.cu).
int launch_stem_rows(fcp_ctx* ctx, const void* src, int mode, int n, int h, int w, Tensor rows);
// conv 3x3/s1 with Cin=3 (RRDB conv_first): f32 NCHW input scaled by in_scale, + bias
int launch_conv3_first(fcp_ctx* ctx, const float* src_nchw, float in_scale, int n, int h, int w, const float* w_kn,
                       const float* shift, Tensor out);
// same, input = u8 NHWC RGB (the value / 255, like rrdb.py:142 on a 0..255 float image)
int launch_conv3_last(fcp_ctx* ctx, Tensor in, const float* packed, Tensor out);
int launch_conv3_first_u8(fcp_ctx* ctx, const uint8_t* src_nhwc, int n, int h, int w, const float* w_kn, const float* shift, Tensor out);
int launch_maxpool3s2(fcp_ctx* ctx, Tensor in, Tensor out);
int launch_global_avgpool(fcp_ctx* ctx, Tensor in, float* out_nc);                     // out[n][c] = mean_hw
// out[n][co] = act((sum_ci in[n][ci] * w[ci][co]) * scale[co] + shift[co])           (1x1 conv on a pooled vector)
int launch_fc(fcp_ctx* ctx, const float* in_nc, int n, int cin, const ConvWeights* wt, int act, float* out_nc);
// out = in * mul[n][c] (+ addvec[n][c]) (+ in) (+ addt nearest-upsampled by up): channel attention / fusion ops
int launch_channel_affine(fcp_ctx* ctx, Tensor in, const float* mul_nc, const float* addvec_nc, int add_self,
                          Tensor out);

// detection post-processing (det_post.cu)
int det_num_priors(int h, int w);
int det_key_capacity(int h, int w);
int launch_heads_to_flat(fcp_ctx* ctx, const float* const* level_ptrs, int n, int h, int w, float* out_heads);
// decode+threshold -> per-image sort+NMS+strategy -> image-ordered append of the kept faces at faces[*face_count...]
// (face_count must be zeroed before the first micro-batch; img_base = batch index of this micro-batch's image 0)
int launch_det_post(fcp_ctx* ctx, const float* const* level_ptrs, const float* heads_flat, int n, int img_base, int h,
                    int w, float vis_thr, float nms_thr, int strategy, int max_faces, float* rec,
                    unsigned long long* keys, unsigned char* supp, int32_t* cand_count, int32_t* kept_count,
                    float* faces, int32_t* face_img, int32_t* face_count);
// faces[f][16] records -> landmarks [f][10] (minus (left, top) of paddings[img]), boxes, scores, anchors
int launch_unpack_faces(fcp_ctx* ctx, const float* faces, const int32_t* face_img, const int32_t* face_count, int cap,
                        const int32_t* paddings, float* landmarks, float* boxes, float* scores, int32_t* anchors);

// RRDBNet.predict gate (rrdb.py:124-141) on the un-padded landmarks; gate[i] = 1: enhance image i
int launch_enhance_gate(fcp_ctx* ctx, const float* landmarks, const int32_t* face_img, const int32_t* face_count, int cap, int n,
                        int h, int w, float threshold, uint8_t* gate);
// [cap + 1][20] float64 metadata records of the multi-GPU all-gather (slot cap: face count)
int launch_pack_records(fcp_ctx* ctx, const float* landmarks, const int32_t* face_img, const double* matrices, const uint8_t* valid,
                        const int32_t* face_count, int cap, int index_base, double* records, cudaStream_t stream);

// align (align.cu)
int launch_solve(fcp_ctx* ctx, const float* landmarks, const int32_t* face_count_dev, int f, const float* target,
                 int allow_skew, double* matrices, double* inv, uint8_t* valid);
int launch_invert(fcp_ctx* ctx, const double* matrices, int f, double* inv);
int launch_warp(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, const uint8_t* const* image_ptrs,
                const int32_t* hs, const int32_t* ws, const int32_t* paddings, const int32_t* indices,
                const int32_t* face_count_dev, int f, const double* inv, const uint8_t* valid, int out_w, int out_h,
                int border, uint8_t* out);

// parse
int launch_parse_prep(fcp_ctx* ctx, const uint8_t* crops, int f, int h, int w, float* out_nhwc3 /*[f,512,512,3]*/);
int launch_parse_tail(fcp_ctx* ctx, const float* logits, int layout_nhwc, int cs, int f, int fh, int fw, int h, int w,
                      uint8_t* labels, int32_t* hist);
int launch_masks(fcp_ctx* ctx, const uint8_t* labels, size_t count, const uint8_t* lut_dev, uint8_t* out);
// grouping rules of bise.py:214-325 on the per-class histogram; all mask groups' 0/255 images in one pass over the labels
int launch_group(fcp_ctx* ctx, const int32_t* hist, int f, const int32_t* codes, const int32_t* offs, int n_attr, int attr_thr,
                 int join_and, const uint8_t* lut, int n_mask, int mask_thr, uint8_t* out_attr, uint8_t* out_mask);
int launch_multi_masks(fcp_ctx* ctx, const uint8_t* labels, size_t count, const uint8_t* lut_dev, int n_mask, uint8_t* out);
int launch_upsample2x(fcp_ctx* ctx, Tensor in, Tensor out);   // out = nearest x2 upsample of in
int launch_nhwc_to_nchw(fcp_ctx* ctx, const float* in, int n, int h, int w, int c, int cs, float* out);

// enhance tail: conv_last output [n,4h,4w,3(+pad)] -> bicubic x0.25 -> clamp*255 round, written NCHW f32 [3,h,w]
int launch_rrdb_tail(fcp_ctx* ctx, Tensor x4, float* out_nchw, int h, int w);
// same, written as u8 NHWC [n,h,w,3] (round(clamp(x,0,1)*255) is integral: the reference's as_numpy cast is exact)
int launch_rrdb_tail_u8(fcp_ctx* ctx, Tensor x4, uint8_t* out_nhwc, int h, int w);

// model graphs (runtime.cu)
int finalize_retinaface(fcp_ctx* ctx);
int finalize_bisenet(fcp_ctx* ctx);
int finalize_rrdbnet(fcp_ctx* ctx);

}  // namespace fcp
