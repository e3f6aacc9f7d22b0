// Shared definitions of the NWD handle between the fp32 (nwd.cu) and tensor-core (nwd_tc.cu) paths.
#pragma once
#include "common.cuh"
#include <vector>

struct cm_nwd {
    float* w_dev = nullptr;      // fp32 packed weights (nwd.cu layout), BN folded
    float* wtc_dev = nullptr;    // tf32-rounded weights of the 7 tensor-core layers in UMMA canonical K-major blocks
    int device = 0;
    int sm_count = 0;
    unsigned char* wmt_dev = nullptr;   // fp16 tap tables + biases of the multi-trace tensor-core kernel (nwd_mt.cu)
    bool mt_ok = true;           // the BN-folded weights fit fp16 (mode 2 is refused otherwise)
    int precision = 0;           // 0 = fp32 CUDA cores, 1 = tf32 tcgen05 (one trace per CTA), 2 = fp16 tcgen05 (multi-trace)
};

namespace cm {
namespace nwdtc {
// packs the tensor-core layers (d2,d3,d4,u1,u2,u3,u4) from the 54 state_dict tensors; returns floats
void pack_tc_weights(const float* const* tensors, std::vector<float>& out);
int launch(cm_nwd* h, const void* in, int in_dtype, void* out, int out_dtype, int K, int monotone_start, double* y,
           double* ss, cudaStream_t st);
}  // namespace nwdtc
namespace nwdmt {
void pack_weights(const float* const* tensors, std::vector<unsigned char>& out);
bool weights_fit_fp16();          // result of the last pack_weights on this thread's call (folded weights within +-65504)
int launch(cm_nwd* h, const void* in, int in_dtype, void* out, int out_dtype, int K, int monotone_start, double* y,
           double* ss, cudaStream_t st);
int debug_cycles(long long* out, int n, int enable);
int debug_dump(void* dev_buf, int stage);
}  // namespace nwdmt
}  // namespace cm
