// Detector magnitude + mismatch loss + residual (one CTA per diffraction pattern, deterministic reduction).
// Reference: adorym/forward_model.py:353 / :371-375 (magnitude = norm or sqrt of summed mode intensities),
// :88-103 (lsq / Poisson mismatch on magnitude or intensity data), :128-136 (beamstop mask).
#include "fft_passes.cuh"
#include "../../include/adorym_b200.h"

__global__ void __launch_bounds__(256) k_detector_loss(int M, int npix, int loss_type, int raw_type, float pm, float inv_n,
                                                       const cpx* det, const float* meas, const float* mask, float* pred,
                                                       cpx* g_det, float* loss_per_pattern) {
#ifndef ADM_HOST_EMU
    __shared__ float red[256];
#else
    float* red = (float*)emu_smem;
#endif
    const int b = blockIdx.x;
    const cpx* d = det + (size_t)b * M * npix;
    float acc = 0.f;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        float s = 0.f;
        for (int m = 0; m < M; ++m) {
            cpx v = d[(size_t)m * npix + i];
            s += v.x * v.x + v.y * v.y;
        }
        float a = sqrtf(s);
        pred[(size_t)b * npix + i] = a;
        bool valid = (mask == nullptr) || (mask[i] >= 1e-5f);
        float y = meas[(size_t)b * npix + i];
        float dlda = 0.f;
        if (valid) {
            if (loss_type == ADM_LOSS_LSQ) {
                float mm = (raw_type == ADM_RAW_MAGNITUDE) ? y : sqrtf(y);
                float r = a - mm;
                acc += r * r;
                dlda = 2.f * r * inv_n;
            } else {
                float Im = (raw_type == ADM_RAW_MAGNITUDE) ? y * y : y;
                float a2 = a * a * pm;
                acc += a2 - Im * pm * logf(a2);
                dlda = (a > 0.f) ? (2.f * pm * a - 2.f * pm * Im / a) * inv_n : 0.f;
            }
        }
        if (g_det != nullptr) {
            float c = (a > 0.f) ? dlda / a : 0.f;      // torch.norm subgradient at 0 (quirk 13)
            for (int m = 0; m < M; ++m) {
                cpx v = d[(size_t)m * npix + i];
                g_det[((size_t)b * M + m) * npix + i] = mk(c * v.x, c * v.y);
            }
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_per_pattern[b] = red[0];
}

extern "C" int adm_detector_loss(int B, int M, int P, int loss_type, int raw_data_type, float poisson_multiplier, float inv_n,
                                 const float* det, const float* meas, const float* mask, float* pred, float* g_det,
                                 float* loss_per_pattern, void* stream) {
    if (B < 1 || M < 1 || P < 1 || !det || !meas || !pred || !loss_per_pattern) return ADM_ERR_ARG;
    if (loss_type != ADM_LOSS_LSQ && loss_type != ADM_LOSS_POISSON) return ADM_ERR_ARG;
    if (raw_data_type != ADM_RAW_MAGNITUDE && raw_data_type != ADM_RAW_INTENSITY) return ADM_ERR_ARG;
    ADM_LAUNCH(k_detector_loss, B, 256, 256 * sizeof(float), stream, M, P * P, loss_type, raw_data_type, poisson_multiplier,
               inv_n, (const cpx*)det, meas, mask, pred, (cpx*)g_det, loss_per_pattern);
    return ADM_OK;
}
