// det_post.cu — RetinaFace post-processing on device: softmax score, analytic priors, SSD decode, threshold
// compaction, per-image sort + greedy IoU NMS, strategy selection, image-ordered output.
//
// Replaces (all host-driven in the reference): PriorBox.forward (_layers.py:49-62), decode_bboxes /
// decode_landms (retinaface.py:146-212) and the x[W,H] scaling (retinaface.py:455-461), filter_preds
// (retinaface.py:214-304), take_by_strategy (retinaface.py:306-408).  HBM-bound integer/float32 work, no tensor
// cores.  Float32 operations use explicit round-to-nearest intrinsics in the reference's operation order so that
// fused multiply-adds cannot change a threshold or IoU decision.
#include "common.h"

namespace fcp {

namespace {

struct DetSrc {
    const float* lvl[3];   // per-level head tensors [n, fh, fw, 32] = (cls 2x2, box 2x4, ldm 2x10); nullptr -> flat
    const float* flat;     // [n, A, 16] = (cls2, box4, ldm10) in prior order
    int fh[3], fw[3], start[3];
    int A;
};

__device__ __forceinline__ void prior_of(const DetSrc& s, int idx, int H, int W, int& lvl, int& cell, int& anc,
                                         float pr[4]) {
    lvl = idx >= s.start[2] ? 2 : (idx >= s.start[1] ? 1 : 0);
    int local = idx - s.start[lvl];
    anc = local & 1;
    cell = local >> 1;
    int i = cell / s.fw[lvl], j = cell - i * s.fw[lvl];
    const double step = lvl == 0 ? 8.0 : (lvl == 1 ? 16.0 : 32.0);
    const double ms = (lvl == 0 ? 16.0 : (lvl == 1 ? 64.0 : 256.0)) * (anc ? 2.0 : 1.0);
    // python-double arithmetic rounded once to float32, like torch.tensor(anchors) (_layers.py:56-62)
    pr[0] = (float)(__ddiv_rn(__dmul_rn((double)j + 0.5, step), (double)W));
    pr[1] = (float)(__ddiv_rn(__dmul_rn((double)i + 0.5, step), (double)H));
    pr[2] = (float)(__ddiv_rn(ms, (double)W));
    pr[3] = (float)(__ddiv_rn(ms, (double)H));
}

__device__ __forceinline__ void fetch_heads(const DetSrc& s, int n, int idx, int lvl, int cell, int anc, float cls[2],
                                            float box[4], float ldm[10]) {
    if (s.flat) {
        const float4* p = reinterpret_cast<const float4*>(s.flat + ((size_t)n * s.A + idx) * 16);
        float4 a = p[0], b = p[1], c = p[2], d = p[3];
        cls[0] = a.x; cls[1] = a.y; box[0] = a.z; box[1] = a.w; box[2] = b.x; box[3] = b.y;
        ldm[0] = b.z; ldm[1] = b.w; ldm[2] = c.x; ldm[3] = c.y; ldm[4] = c.z; ldm[5] = c.w;
        ldm[6] = d.x; ldm[7] = d.y; ldm[8] = d.z; ldm[9] = d.w;
    } else {
        const float* p = s.lvl[lvl] + ((size_t)n * s.fh[lvl] * s.fw[lvl] + cell) * 32;
        cls[0] = p[2 * anc]; cls[1] = p[2 * anc + 1];
#pragma unroll
        for (int k = 0; k < 4; ++k) box[k] = p[4 + 4 * anc + k];
#pragma unroll
        for (int k = 0; k < 10; ++k) ldm[k] = p[12 + 10 * anc + k];
    }
}

__device__ __forceinline__ float face_score(const float cls[2]) {
    // softmax(dim=-1)[..., 1] (retinaface.py:144,458)
    float m = fmaxf(cls[0], cls[1]);
    float e0 = expf(__fsub_rn(cls[0], m)), e1 = expf(__fsub_rn(cls[1], m));
    return __fdiv_rn(e1, __fadd_rn(e0, e1));
}

// one thread per (image, prior): score -> threshold -> decode -> record at rec[n][prior], key appended to keys[n]
__global__ void det_decode_kernel(DetSrc s, int N, int H, int W, float vis_thr, float* __restrict__ rec,
                                  unsigned long long* __restrict__ keys, int key_cap, int* __restrict__ count) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int n = blockIdx.y;
    if (idx >= s.A) return;
    int lvl, cell, anc;
    float pr[4], cls[2], box[4], ldm[10];
    prior_of(s, idx, H, W, lvl, cell, anc, pr);
    fetch_heads(s, n, idx, lvl, cell, anc, cls, box, ldm);
    float score = face_score(cls);
    if (!(score > vis_thr)) return;                                   // masks = scores > vis_threshold (retinaface.py:264)
    const float fw = (float)W, fh = (float)H;
    float cx = __fadd_rn(pr[0], __fmul_rn(__fmul_rn(box[0], 0.1f), pr[2]));
    float cy = __fadd_rn(pr[1], __fmul_rn(__fmul_rn(box[1], 0.1f), pr[3]));
    float bw = __fmul_rn(pr[2], expf(__fmul_rn(box[2], 0.2f)));
    float bh = __fmul_rn(pr[3], expf(__fmul_rn(box[3], 0.2f)));
    float x1 = __fsub_rn(cx, __fdiv_rn(bw, 2.f)), y1 = __fsub_rn(cy, __fdiv_rn(bh, 2.f));   // boxes[..., :2] -= boxes[..., 2:] / 2
    float x2 = __fadd_rn(bw, x1), y2 = __fadd_rn(bh, y1);                                   // boxes[..., 2:] += boxes[..., :2]
    float* r = rec + ((size_t)n * s.A + idx) * 16;
    float4 o0, o1, o2, o3;
    o0.x = score; o0.y = __int_as_float(idx);
    o0.z = __fmul_rn(x1, fw); o0.w = __fmul_rn(y1, fh);
    o1.x = __fmul_rn(x2, fw); o1.y = __fmul_rn(y2, fh);
    float l[10];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        l[2 * k] = __fmul_rn(__fadd_rn(pr[0], __fmul_rn(__fmul_rn(ldm[2 * k], 0.1f), pr[2])), fw);
        l[2 * k + 1] = __fmul_rn(__fadd_rn(pr[1], __fmul_rn(__fmul_rn(ldm[2 * k + 1], 0.1f), pr[3])), fh);
    }
    o1.z = l[0]; o1.w = l[1];
    o2 = make_float4(l[2], l[3], l[4], l[5]);
    o3 = make_float4(l[6], l[7], l[8], l[9]);
    float4* r4 = reinterpret_cast<float4*>(r);
    r4[0] = o0; r4[1] = o1; r4[2] = o2; r4[3] = o3;
    int slot = atomicAdd(&count[n], 1);
    // ascending key order == descending score, ties broken by ascending prior index
    unsigned long long key = ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(score)) << 32) | (unsigned)idx;
    if (slot < key_cap) keys[(size_t)n * key_cap + slot] = key;
}

__device__ __forceinline__ float box_area(const float4& b) {
    return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

// one CTA per image: bitonic sort of the candidate keys, greedy NMS, strategy.  Kept priors are written to the
// front of keys[n] (low 32 bits) and their number to kept_count[n].
constexpr int NMS_THREADS = 1024;
constexpr int NMS_SMEM_KEYS = 4096;

__global__ void __launch_bounds__(NMS_THREADS) det_nms_kernel(const float* __restrict__ rec,
                                                              unsigned long long* __restrict__ keys_all, int key_cap,
                                                              const int* __restrict__ count, int A, float nms_thr,
                                                              int strategy, unsigned char* __restrict__ supp_all,
                                                              int* __restrict__ kept_count) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    float4* sbox = reinterpret_cast<float4*>(nms_smem);                                         // [NMS_SMEM_KEYS]
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sbox + NMS_SMEM_KEYS);    // [NMS_SMEM_KEYS]
    unsigned char* ssupp = reinterpret_cast<unsigned char*>(skeys + NMS_SMEM_KEYS);             // [NMS_SMEM_KEYS]
    const int n = blockIdx.x, tid = threadIdx.x;
    const int K = min(count[n], key_cap);
    unsigned long long* gkeys = keys_all + (size_t)n * key_cap;
    const float* grec = rec + (size_t)n * A * 16;
    if (K == 0) {
        if (tid == 0) kept_count[n] = 0;
        return;
    }
    int P = 1;
    while (P < K) P <<= 1;
    const bool in_smem = P <= NMS_SMEM_KEYS;
    unsigned long long* keys = in_smem ? skeys : gkeys;
    unsigned char* supp = in_smem ? ssupp : supp_all + (size_t)n * key_cap;
    for (int i = tid; i < P; i += NMS_THREADS) {
        unsigned long long v = i < K ? gkeys[i] : ~0ull;
        if (in_smem) skeys[i] = v; else if (i >= K) gkeys[i] = v;
        supp[i] = 0;
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += NMS_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = keys[i], b = keys[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    if (in_smem) {
        for (int i = tid; i < K; i += NMS_THREADS) {
            int anchor = (int)(skeys[i] & 0xFFFFFFFFu);
            const float4* r4 = reinterpret_cast<const float4*>(grec + (size_t)anchor * 16);
            float4 a = r4[0], b = r4[1];
            sbox[i] = make_float4(a.z, a.w, b.x, b.y);
        }
    }
    __syncthreads();
    auto box_at = [&](int i) -> float4 {
        if (in_smem) return sbox[i];
        int anchor = (int)(keys[i] & 0xFFFFFFFFu);
        const float4* r4 = reinterpret_cast<const float4*>(grec + (size_t)anchor * 16);
        float4 a = r4[0], b = r4[1];
        return make_float4(a.z, a.w, b.x, b.y);
    };
    int nkept = 0;
    for (int i = 0; i < K; ++i) {
        if (supp[i]) continue;                               // uniform: flags of index i are final here
        float4 bi = box_at(i);
        float ai = box_area(bi);
        unsigned anchor_i = (unsigned)(keys[i] & 0xFFFFFFFFu);
        __syncthreads();                                      // everyone has read keys[i] before slot nkept is reused
        if (tid == 0) gkeys[nkept] = anchor_i;
        ++nkept;
        if (strategy == FCP_STRATEGY_BEST) break;
        for (int j = i + 1 + tid; j < K; j += NMS_THREADS) {
            if (supp[j]) continue;
            float4 bj = box_at(j);
            float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
            float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
            float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
            float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
            float inter = __fmul_rn(w, h);
            float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, box_area(bj)), inter));
            if (!(ovr <= nms_thr)) supp[j] = 1;              // keep only ovr <= nms_threshold (retinaface.py:291)
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (strategy == FCP_STRATEGY_LARGEST && nkept > 0) {
            // argmax of the (+1) area among kept, first maximum wins (retinaface.py:390-398)
            int best = 0;
            float best_a = -INFINITY;
            for (int p = 0; p < nkept; ++p) {
                const float4* r4 = reinterpret_cast<const float4*>(grec + (size_t)(unsigned)gkeys[p] * 16);
                float4 a = r4[0], b = r4[1];
                float ar = box_area(make_float4(a.z, a.w, b.x, b.y));
                if (ar > best_a) { best_a = ar; best = p; }
            }
            gkeys[0] = gkeys[best];
            nkept = 1;
        }
        kept_count[n] = nkept;
    }
}

// single CTA: exclusive scan of kept counts over images, then image-ordered copy of the face records
__global__ void __launch_bounds__(1024) det_gather_kernel(const float* __restrict__ rec,
                                                          const unsigned long long* __restrict__ keys_all, int key_cap,
                                                          const int* __restrict__ kept_count, int N, int A,
                                                          int img_base, int max_faces, float* __restrict__ faces,
                                                          int* __restrict__ face_img, int* __restrict__ face_count) {
    extern __shared__ int offs[];   // [N+1]
    if (threadIdx.x == 0) {
        int acc = *face_count;      // faces of the previous micro-batches (stream order makes this safe)
        for (int n = 0; n < N; ++n) { offs[n] = acc; acc += kept_count[n]; }
        offs[N] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) *face_count = offs[N];
    for (int n = 0; n < N; ++n) {
        int cnt = offs[n + 1] - offs[n];
        for (int p = threadIdx.x; p < cnt; p += blockDim.x) {
            int dst = offs[n] + p;
            if (dst >= max_faces) continue;
            unsigned anchor = (unsigned)keys_all[(size_t)n * key_cap + p];
            const float4* r4 = reinterpret_cast<const float4*>(rec + ((size_t)n * A + anchor) * 16);
            float4* d4 = reinterpret_cast<float4*>(faces + (size_t)dst * 16);
            d4[0] = r4[0]; d4[1] = r4[1]; d4[2] = r4[2]; d4[3] = r4[3];
            face_img[dst] = img_base + n;
        }
    }
}

// head tensors [n,fh,fw,32] per level -> flat [n,A,16] in the reference's prior order (test/diagnostic output)
__global__ void heads_to_flat_kernel(DetSrc s, float* __restrict__ out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int n = blockIdx.y;
    if (idx >= s.A) return;
    int lvl = idx >= s.start[2] ? 2 : (idx >= s.start[1] ? 1 : 0);
    int local = idx - s.start[lvl];
    int anc = local & 1, cell = local >> 1;
    float cls[2], box[4], ldm[10];
    fetch_heads(s, n, idx, lvl, cell, anc, cls, box, ldm);
    float* o = out + ((size_t)n * s.A + idx) * 16;
    o[0] = cls[0]; o[1] = cls[1];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[2 + k] = box[k];
#pragma unroll
    for (int k = 0; k < 10; ++k) o[6 + k] = ldm[k];
}

// faces[f][16] -> landmarks (x - left, y - top of the image's padding: cropper.py:822), boxes, scores, prior ids
__global__ void unpack_faces_kernel(const float* __restrict__ faces, const int* __restrict__ face_img,
                                    const int* __restrict__ face_count, int cap, const int* __restrict__ paddings,
                                    float* __restrict__ lms, float* __restrict__ boxes, float* __restrict__ scores,
                                    int* __restrict__ anchors) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap || i >= *face_count) return;
    const float* r = faces + (size_t)i * 16;
    float left = 0.f, top = 0.f;
    if (paddings) { top = (float)paddings[face_img[i] * 4]; left = (float)paddings[face_img[i] * 4 + 2]; }
    if (lms)
        for (int k = 0; k < 5; ++k) {
            lms[i * 10 + 2 * k] = __fsub_rn(r[6 + 2 * k], left);
            lms[i * 10 + 2 * k + 1] = __fsub_rn(r[7 + 2 * k], top);
        }
    if (boxes) for (int k = 0; k < 4; ++k) boxes[i * 4 + k] = r[2 + k];
    if (scores) scores[i] = r[0];
    if (anchors) anchors[i] = __float_as_int(r[1]);
}

// numpy's pairwise summation of a float32 vector (loops_utils.h.src, PW_BLOCKSIZE 128): what `ndarray.mean()` of
// rrdb.py:141 accumulates with.  v(i) returns element i.
template <class F>
__device__ float np_pairwise_sum(F v, int lo, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, v(lo + i));
        return res;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = v(lo + j);
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], v(lo + i + j));
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, v(lo + i));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum(v, lo, n2), np_pairwise_sum(v, lo + n2, n - n2));
}

// RRDBNet.predict's per-image gate (rrdb.py:124-141): enhance image i iff mean_f((x4-x0)*(y4-y0) / (H*W)) <= threshold over
// its faces, all in float32 like numpy; images without faces are skipped.  faces are stored in image order.
__global__ void enhance_gate_kernel(const float* __restrict__ lms, const int* __restrict__ face_img,
                                    const int* __restrict__ face_count, int cap, int n, float hw, float thr,
                                    unsigned char* __restrict__ gate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int f = min(*face_count, cap);
    int a = 0, b = f;                                     // [lo, hi) = faces of image i (face_img is ascending)
    while (a < b) { const int m = (a + b) >> 1; if (face_img[m] < i) a = m + 1; else b = m; }
    const int lo = a;
    b = f;
    while (a < b) { const int m = (a + b) >> 1; if (face_img[m] <= i) a = m + 1; else b = m; }
    const int hi = a;
    if (hi == lo) { gate[i] = 0; return; }                // no landmarks found: the image is skipped (rrdb.py:133-135)
    auto factor = [&](int k) {
        const float* l = lms + (size_t)k * 10;
        const float w = __fsub_rn(l[8], l[0]), h = __fsub_rn(l[9], l[1]);
        return __fdiv_rn(__fmul_rn(w, h), hw);
    };
    const int cnt = hi - lo;
    const float mean = __fdiv_rn(np_pairwise_sum(factor, lo, cnt), (float)cnt);
    gate[i] = mean <= thr ? 1 : 0;
}

// per-face metadata record of the all-gather (distributed.py RECORD = 20 float64): landmarks[10], GLOBAL image index,
// matrix[6], valid, 2 reserved; slot `cap` carries the face count in column 0
__global__ void pack_records_kernel(const float* __restrict__ lms, const int* __restrict__ face_img, const double* __restrict__ mats,
                                    const unsigned char* __restrict__ valid, const int* __restrict__ face_count, int cap,
                                    int index_base, double* __restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > cap) return;
    double* r = rec + (size_t)i * 20;
    const int f = min(*face_count, cap);
    if (i == cap) { r[0] = (double)f; for (int k = 1; k < 20; ++k) r[k] = 0.0; return; }
    if (i >= f) { for (int k = 0; k < 20; ++k) r[k] = 0.0; return; }
    for (int k = 0; k < 10; ++k) r[k] = (double)lms[(size_t)i * 10 + k];
    r[10] = (double)(face_img[i] + index_base);
    for (int k = 0; k < 6; ++k) r[11 + k] = mats ? mats[(size_t)i * 6 + k] : 0.0;
    r[17] = valid ? (double)valid[i] : 1.0;
    r[18] = r[19] = 0.0;
}

DetSrc make_src(const float* const* level_ptrs, const float* flat, int h, int w) {
    DetSrc s{};
    int acc = 0;
    const int steps[3] = {8, 16, 32};
    for (int l = 0; l < 3; ++l) {
        s.fh[l] = (h + steps[l] - 1) / steps[l];
        s.fw[l] = (w + steps[l] - 1) / steps[l];
        s.start[l] = acc;
        acc += s.fh[l] * s.fw[l] * 2;
        s.lvl[l] = level_ptrs ? level_ptrs[l] : nullptr;
    }
    s.A = acc;
    s.flat = flat;
    return s;
}

}  // namespace

int det_num_priors(int h, int w) {
    return make_src(nullptr, nullptr, h, w).A;
}

int det_key_capacity(int h, int w) {
    int a = det_num_priors(h, w), p = 1;
    while (p < a) p <<= 1;
    return p;
}

int launch_heads_to_flat(fcp_ctx* ctx, const float* const* level_ptrs, int n, int h, int w, float* out_heads) {
    DetSrc s = make_src(level_ptrs, nullptr, h, w);
    dim3 grid((s.A + 255) / 256, n);
    heads_to_flat_kernel<<<grid, 256, 0, ctx->stream>>>(s, out_heads);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_unpack_faces(fcp_ctx* ctx, const float* faces, const int32_t* face_img, const int32_t* face_count, int cap,
                        const int32_t* paddings, float* landmarks, float* boxes, float* scores, int32_t* anchors) {
    if (cap == 0) return FCP_OK;
    unpack_faces_kernel<<<(cap + 127) / 128, 128, 0, ctx->stream>>>(faces, face_img, face_count, cap, paddings, landmarks,
                                                                  boxes, scores, anchors);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_enhance_gate(fcp_ctx* ctx, const float* landmarks, const int32_t* face_img, const int32_t* face_count, int cap, int n,
                        int h, int w, float threshold, uint8_t* gate) {
    if (n == 0) return FCP_OK;
    enhance_gate_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(landmarks, face_img, face_count, cap, n, (float)(h * w), threshold, gate);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_pack_records(fcp_ctx* ctx, const float* landmarks, const int32_t* face_img, const double* matrices, const uint8_t* valid,
                        const int32_t* face_count, int cap, int index_base, double* records, cudaStream_t stream) {
    pack_records_kernel<<<(cap + 1 + 127) / 128, 128, 0, stream>>>(landmarks, face_img, matrices, valid, face_count, cap, index_base, records);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_det_post(fcp_ctx* ctx, const float* const* level_ptrs, const float* heads_flat, int n, int img_base, int h,
                    int w, float vis_thr, float nms_thr, int strategy, int max_faces, float* rec,
                    unsigned long long* keys, unsigned char* supp, int32_t* cand_count, int32_t* kept_count,
                    float* faces, int32_t* face_img, int32_t* face_count) {
    DetSrc s = make_src(level_ptrs, heads_flat, h, w);
    int key_cap = det_key_capacity(h, w);
    FCP_CUDA(ctx, cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * n, ctx->stream));
    dim3 grid((s.A + 255) / 256, n);
    det_decode_kernel<<<grid, 256, 0, ctx->stream>>>(s, n, h, w, vis_thr, rec, keys, key_cap, cand_count);
    FCP_KERNEL_CHECK(ctx);
    const size_t nms_smem = (size_t)NMS_SMEM_KEYS * (sizeof(float4) + sizeof(unsigned long long) + 1);
    static uint64_t configured = 0;                    // one bit per device: the attribute is per device
    if (!((configured >> (ctx->device & 63)) & 1)) {
        FCP_CUDA(ctx, cudaFuncSetAttribute(det_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem));
        configured |= (uint64_t)1 << (ctx->device & 63);
    }
    det_nms_kernel<<<n, NMS_THREADS, nms_smem, ctx->stream>>>(rec, keys, key_cap, cand_count, s.A, nms_thr, strategy, supp, kept_count);
    FCP_KERNEL_CHECK(ctx);
    det_gather_kernel<<<1, 1024, sizeof(int) * (n + 1), ctx->stream>>>(rec, keys, key_cap, kept_count, n, s.A, img_base, max_faces,
                                                                      faces, face_img, face_count);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

}  // namespace fcp
