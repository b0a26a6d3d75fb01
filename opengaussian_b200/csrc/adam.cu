// adam.cu -- the optimiser step of all parameter groups in ONE launch.
//
// Replaces `gaussians.optimizer.step()` (train.py:609; torch.optim.Adam(l, lr=0.0, eps=1e-15) built at
// scene/gaussian_model.py:215-230 with seven groups: xyz, f_dc, f_rest, opacity, scaling, rotation, ins_feat).
// torch's default multi-tensor path makes ~10 elementwise passes over the 59 floats per Gaussian; here every
// element is read once (grad, param, exp_avg, exp_avg_sq: 16 B) and written once (12 B): 28 B per element,
// HBM-bound.  The arithmetic follows torch/optim/adam.py::_single_tensor_adam (no weight decay, no amsgrad):
//     exp_avg    <- exp_avg + (1 - beta1) (g - exp_avg)                      (lerp_)
//     exp_avg_sq <- beta2 exp_avg_sq + (1 - beta2) g g                       (mul_, addcmul_)
//     denom      <- sqrt(exp_avg_sq) / sqrt(1 - beta2^t) + eps
//     param      <- param - (lr / (1 - beta1^t)) exp_avg / denom             (addcdiv_)
// with IEEE sqrt and division.  The per-tensor scalars (lr / (1 - beta1^t), sqrt(1 - beta2^t)) come from the host.
#include "common.cuh"

namespace ogs {

#define AD_THREADS 256
#define AD_CTA_ELEMS 2048          // 2 x float4 per thread
#define AD_MAX_TENSORS 16

static_assert(sizeof(ogs_adam_tensor) == 72, "ogs_adam_tensor layout is part of the C ABI (ctypes mirror in _lib.py)");

struct AdamLaunch {
    ogs_adam_tensor t[AD_MAX_TENSORS];
    int first_block[AD_MAX_TENSORS + 1];
    int vec[AD_MAX_TENSORS];       // all four pointers 16-byte aligned
    int count;
    float grad_scale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const ogs_adam_tensor& T, float gs) {
    g *= gs;
    m = fmaf(T.one_minus_beta1, g - m, m);
    v = fmaf(v, T.beta2, T.one_minus_beta2 * g * g);
    const float denom = __fdiv_rn(__fsqrt_rn(v), T.bias_correction2_sqrt) + T.eps;
    p = fmaf(-T.step_size, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(AD_THREADS) adam_kernel(const __grid_constant__ AdamLaunch L) {
    const int b = blockIdx.x;
    int k = 0;
    while (k + 1 < L.count && b >= L.first_block[k + 1]) ++k;
    const ogs_adam_tensor& T = L.t[k];
    const int64_t base = (int64_t)(b - L.first_block[k]) * AD_CTA_ELEMS;
    const float gs = L.grad_scale;
#pragma unroll
    for (int u = 0; u < AD_CTA_ELEMS / (AD_THREADS * 4); ++u) {
        const int64_t i = base + ((int64_t)u * AD_THREADS + threadIdx.x) * 4;
        if (i >= T.n) continue;
        if (L.vec[k] && i + 4 <= T.n) {
            const float4 g = __ldcs(reinterpret_cast<const float4*>(T.grad + i));
            float4 p = *reinterpret_cast<const float4*>(T.param + i);
            float4 m = *reinterpret_cast<const float4*>(T.exp_avg + i);
            float4 v = *reinterpret_cast<const float4*>(T.exp_avg_sq + i);
            adam_one(p.x, g.x, m.x, v.x, T, gs);
            adam_one(p.y, g.y, m.y, v.y, T, gs);
            adam_one(p.z, g.z, m.z, v.z, T, gs);
            adam_one(p.w, g.w, m.w, v.w, T, gs);
            *reinterpret_cast<float4*>(T.param + i) = p;
            *reinterpret_cast<float4*>(T.exp_avg + i) = m;
            *reinterpret_cast<float4*>(T.exp_avg_sq + i) = v;
        } else {
            const int64_t e = (i + 4 < T.n) ? i + 4 : T.n;
            for (int64_t q = i; q < e; ++q) {
                float p = T.param[q], m = T.exp_avg[q], v = T.exp_avg_sq[q];
                adam_one(p, T.grad[q], m, v, T, gs);
                T.param[q] = p;
                T.exp_avg[q] = m;
                T.exp_avg_sq[q] = v;
            }
        }
    }
}

int launch_adam_step(int n_tensors, const ogs_adam_tensor* tensors, float grad_scale, cudaStream_t s) {
    int done = 0;
    while (done < n_tensors) {
        AdamLaunch L;
        L.count = 0;
        L.grad_scale = grad_scale;
        int64_t blocks = 0;
        while (done < n_tensors && L.count < AD_MAX_TENSORS) {
            const ogs_adam_tensor& t = tensors[done];
            if (t.n < 0 || (t.n > 0 && (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq))) {
                set_error("adam_step: tensor %d has a null pointer or a negative size", done);
                return -1;
            }
            const int64_t nb = (t.n + AD_CTA_ELEMS - 1) / AD_CTA_ELEMS;
            if (blocks + nb > 0x7FFFFFFF) {
                if (L.count == 0) { set_error("adam_step: tensor %d is too large for one launch", done); return -1; }
                break;
            }
            ++done;
            if (t.n == 0) continue;
            L.t[L.count] = t;
            L.first_block[L.count] = (int)blocks;
            L.vec[L.count] = ((((uintptr_t)t.param | (uintptr_t)t.grad | (uintptr_t)t.exp_avg | (uintptr_t)t.exp_avg_sq) & 15) == 0);
            blocks += nb;
            ++L.count;
        }
        L.first_block[L.count] = (int)blocks;
        if (L.count == 0) continue;
        adam_kernel<<<(unsigned)blocks, AD_THREADS, 0, s>>>(L);
    }
    return 0;
}

}  // namespace ogs
