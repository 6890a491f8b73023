// STAGED EXPERIMENT -- not compiled into the library (see experiments/README.md).
//
// TRAIN.OPTIMIZER = 'ADAM' (the third optimiser BiaPy's configuration accepts next to SGD and ADAMW; the engine raises
// NotImplementedError for it today).  timm's create_optimizer_v2('adam', weight_decay=wd) is torch.optim.Adam: the decay is an L2
// term added to the gradient before the moments, not the decoupled shrink of AdamW.  Same thread layout, moment update and
// bias-correction arguments as the product's adamw_kernel (ops.cu); `decoupled` = 1 reproduces it exactly.
#pragma once

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale,
                            int decoupled) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * grad_scale;
    if (decoupled) pi *= (1.f - lr * wd);                // AdamW: p <- p (1 - lr wd)
    else gi = fmaf(wd, pi, gi);                          // Adam:  g <- g + wd p
    const float mi = m[i] + (1.f - b1) * (gi - m[i]);
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}
