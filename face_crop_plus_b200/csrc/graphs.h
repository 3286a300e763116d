// graphs.h — helpers shared by runtime.cu and graphs.cu: staging of host/device buffers, graph executor.
#pragma once
#include "common.h"

namespace fcp {

bool is_device_ptr(const void* p);
int pack_conv(fcp_ctx* ctx, Model& m, const std::vector<std::string>& convs, const std::string& bn,
              const std::string& name, const float* explicit_scale = nullptr, const float* explicit_shift = nullptr);

// tensor-core packing of a 7x7 stem (`conv` = name of the already packed square conv, whose folded BN it shares):
// weights [64][7 vertical taps][32 = 7 horizontal taps x 3 channels, zero padded], registered as `name`
int pack_stem_rows(fcp_ctx* ctx, Model& m, const std::string& conv, const std::string& name);
// tensor-core packing of the detector stem for the direct uint8 route (no row-patch tensor): [64][4 tap-row pairs][64]
int pack_stem_direct(fcp_ctx* ctx, Model& m, const std::string& conv, const std::string& name);

// borrowed input: used in place when it already lives on the device, otherwise copied H2D on the context stream
class DevIn {
public:
    int init(fcp_ctx* ctx, const void* src, size_t bytes);
    ~DevIn();
    template <class T> const T* as() const { return static_cast<const T*>(p_); }
private:
    fcp_ctx* ctx_ = nullptr;
    void* p_ = nullptr;
    bool owned_ = false;
};

// caller-allocated output: written in place when on the device, otherwise produced in a scratch buffer and copied D2H
// by flush().  need_scratch: allocate a device buffer even when dst == nullptr (the value is needed internally).
class DevOut {
public:
    int init(fcp_ctx* ctx, void* dst, size_t bytes, bool need_scratch = false);
    int flush(size_t bytes = (size_t)-1);
    ~DevOut();
    template <class T> T* as() const { return static_cast<T*>(p_); }
private:
    fcp_ctx* ctx_ = nullptr;
    void* dst_ = nullptr;
    void* p_ = nullptr;
    size_t bytes_ = 0;
    bool owned_ = false;
};

// Graph executor: allocates activations from ctx->arena and launches kernels, or only plans (arena in plan mode).
struct Exec {
    fcp_ctx* ctx;
    Model* model;
    bool dry;
    int status = FCP_OK;
    bool ok() const { return status == FCP_OK; }
    Tensor alloc(int n, int h, int w, int c, int cs = 0);
    float* alloc_vec(size_t count);
    void free(Tensor& t);
    void free_vec(float* p);
    const ConvWeights* W(const std::string& name);
    // out = epilogue(conv(in)); returns false on error (status set)
    bool conv(const std::string& name, Tensor in, Tensor out, int stride, int pad, int act, ConvOp extra = ConvOp());
};

}  // namespace fcp
