// Internal interface of the fused strict UMNN kernels (tc_umnn3.cu) for umnn_lw.cu.
#pragma once
#include "common.cuh"

#ifndef GNF_EMU
namespace gnf {
// floats of scratch for the packed W_l^T chunk images of the backward chain (0: integrand not covered, see gnf_last_error)
size_t u3_bwd_image_floats(const gnf_mlp_t* net);
// delta_L .. delta_1 in one kernel; writes delta_{L-1} .. delta_2 into dplanes, accumulates db (hidden), dW0[:,0], D, dx atomically
int launch_u3_bwd_chain(const float* x, const gnf_mlp_t* net, int S, const float* ccw, const float* ccn, const float* jac, const float* gz,
                        const float* gzrev, const float* gjac, const float* glogdet, const float* saved, float* image, float* dplanes, float* D,
                        float* dx, const gnf_mlp_grad_t* grads, int R, int d, cudaStream_t s, const Branches* br = nullptr, int side = 0,
                        cudaStream_t pack_stream = nullptr);
}  // namespace gnf
#endif
