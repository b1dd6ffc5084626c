// Sub-pixel Fourier shift of the probe, one CTA per (pattern, mode), and its transpose (probe and shift gradients).
// Reference: adorym/util.py:380-397 realign_image_fourier as called per position at forward_model.py:296-303:
//   out = IFFT2( FFT2(probe) * exp(-2 pi i (fx * shift_x + fy * shift_y)) ),  fx, fy = fftfreq(P), shift = (y, x).
#include "fft_passes.cuh"
#include "../../include/adorym_b200.h"

#define ADM_PI_REF 3.14159265359f   // adorym/constants.py

template <class C> ADM_DEV float fftfreq_of(int k) { return (float)(k < C::P / 2 ? k : k - C::P) / (float)C::P; }

template <class C> ADM_DEV cpx ramp_at(int py_pos, int px_pos, float sy, float sx) {
    float fy = fftfreq_of<C>(perm_freq<C>(py_pos));
    float fx = fftfreq_of<C>(perm_freq<C>(px_pos));
    float ph = -2.f * ADM_PI_REF * (fx * sx + fy * sy);
    float s, c;
#ifndef ADM_HOST_EMU
    sincosf(ph, &s, &c);
#else
    s = sinf(ph); c = cosf(ph);
#endif
    return mk(c, s);
}

// load a natural-order [P][P] image into the padded smem buffer (lanes along x)
template <class C> ADM_DEV void load_natural(cpx* buf, const cpx* src, int q, int l) {
    for (int y = q; y < C::P; y += C::Q) buf[y * C::S + l] = src[y * C::P + l];
}
// full forward 2-D FFT in smem (natural in, permuted spectrum out)
template <class C> ADM_DEV void fft2_fwd(cpx* buf, const cpx* tw, int q, int l, bool active) {
    if (active) pass_A_fwd<C, false>(buf, tw, q, l);
    __syncthreads();
    if (active) pass_B_fwd<C, false>(buf, q, l);
    __syncthreads();
    if (active) pass_A_fwd<C, true>(buf, tw, q, l);
    __syncthreads();
    if (active) pass_B_fwd<C, true>(buf, q, l);
    __syncthreads();
}
// full unnormalised inverse 2-D FFT in smem (permuted spectrum in, natural out)
template <class C> ADM_DEV void fft2_inv(cpx* buf, const cpx* tw, int q, int l, bool active) {
    if (active) pass_B_inv<C, true>(buf, tw, q, l);
    __syncthreads();
    if (active) pass_A_inv<C, true>(buf, q, l);
    __syncthreads();
    if (active) pass_B_inv<C, false>(buf, tw, q, l);
    __syncthreads();
    if (active) pass_A_inv<C, false>(buf, q, l);
    __syncthreads();
}

template <class C> __global__ void __launch_bounds__(C::NTHREADS) k_probe_shift_fwd(int M, const cpx* probe, const float* shift, cpx* out) {
    ADM_SMEM_DECL
    constexpr int P = C::P, S = C::S;
    cpx* buf = (cpx*)smem_raw;
    cpx* tw = buf + P * S;
    const int t = threadIdx.x;
    const bool active = t < C::NACT;
    const int q = t / P, l = t % P;
    const int wf = blockIdx.x, b = wf / M, m = wf % M;
    const float sy = shift[2 * b], sx = shift[2 * b + 1];
    init_twiddles<C>(tw, t);
    if (active) load_natural<C>(buf, probe + (size_t)m * P * P, q, l);
    __syncthreads();
    fft2_fwd<C>(buf, tw, q, l, active);
    const float sc = 1.0f / (float)(P * P);
    if (active)
        for (int y = q; y < P; y += C::Q) buf[y * S + l] = cscale(cmul(buf[y * S + l], ramp_at<C>(y, l, sy, sx)), sc);
    __syncthreads();
    fft2_inv<C>(buf, tw, q, l, active);
    if (active)
        for (int y = q; y < P; y += C::Q) out[(size_t)wf * P * P + y * P + l] = buf[y * S + l];
}

template <class C> __global__ void __launch_bounds__(C::NTHREADS) k_probe_shift_bwd(int M, const cpx* probe, const float* shift,
                                                                                     const cpx* g_out, cpx* g_probe, float* g_shift) {
    ADM_SMEM_DECL
    constexpr int P = C::P, S = C::S;
    cpx* bufp = (cpx*)smem_raw;          // spectrum of the probe
    cpx* bufg = bufp + P * S;            // spectrum of the incoming gradient
    cpx* tw = bufg + P * S;
    float* red = (float*)(tw + P);       // [2][NTHREADS]
    const int t = threadIdx.x;
    const bool active = t < C::NACT;
    const int q = t / P, l = t % P;
    const int wf = blockIdx.x, b = wf / M, m = wf % M;
    const float sy = shift[2 * b], sx = shift[2 * b + 1];
    init_twiddles<C>(tw, t);
    if (active) {
        load_natural<C>(bufp, probe + (size_t)m * P * P, q, l);
        load_natural<C>(bufg, g_out + (size_t)wf * P * P, q, l);
    }
    __syncthreads();
    fft2_fwd<C>(bufp, tw, q, l, active);
    fft2_fwd<C>(bufg, tw, q, l, active);
    // dL/ds = (2 pi / P^2) sum_k f_k Im(conj(G_k) S_k),  S = F(probe) * ramp ; g_probe = IFFT2(conj(ramp) G)
    float ay = 0.f, ax = 0.f;
    const float sc = 1.0f / (float)(P * P);
    if (active)
        for (int y = q; y < P; y += C::Q) {
            cpx r = ramp_at<C>(y, l, sy, sx);
            cpx Sk = cmul(bufp[y * S + l], r);
            cpx G = bufg[y * S + l];
            float im = G.x * Sk.y - G.y * Sk.x;            // Im(conj(G) S)
            ay += fftfreq_of<C>(perm_freq<C>(y)) * im;
            ax += fftfreq_of<C>(perm_freq<C>(l)) * im;
            bufg[y * S + l] = cscale(cmulc(G, r), sc);
        }
    red[t] = ay;
    red[C::NTHREADS + t] = ax;
    __syncthreads();
    if (t == 0) {
        float sy_ = 0.f, sx_ = 0.f;
        for (int i = 0; i < C::NTHREADS; ++i) { sy_ += red[i]; sx_ += red[C::NTHREADS + i]; }
        const float c = 2.f * ADM_PI_REF * sc;
        atomicAdd(&g_shift[2 * b], c * sy_);
        atomicAdd(&g_shift[2 * b + 1], c * sx_);
    }
    fft2_inv<C>(bufg, tw, q, l, active);
    if (active && g_probe != nullptr)
        for (int y = q; y < P; y += C::Q) {
            cpx v = bufg[y * S + l];
            atomicAdd(&g_probe[(size_t)m * P * P + y * P + l].x, v.x);
            atomicAdd(&g_probe[(size_t)m * P * P + y * P + l].y, v.y);
        }
}

template <class C> static int launch_shift_fwd(int B, int M, const float* probe, const float* shift, float* out, void* stream) {
    size_t smem = sizeof(cpx) * (C::P * C::S + C::P);
    if (set_smem(k_probe_shift_fwd<C>, smem)) return ADM_ERR_CUDA;
    ADM_LAUNCH(k_probe_shift_fwd<C>, B * M, C::NTHREADS, smem, stream, M, (const cpx*)probe, shift, (cpx*)out);
    return ADM_OK;
}
template <class C> static int launch_shift_bwd(int B, int M, const float* probe, const float* shift, const float* g_out,
                                               float* g_probe, float* g_shift, void* stream) {
    size_t smem = sizeof(cpx) * (2 * C::P * C::S + C::P) + sizeof(float) * 2 * C::NTHREADS;
    if (set_smem(k_probe_shift_bwd<C>, smem)) return ADM_ERR_CUDA;
    ADM_LAUNCH(k_probe_shift_bwd<C>, B * M, C::NTHREADS, smem, stream, M, (const cpx*)probe, shift, (const cpx*)g_out,
               (cpx*)g_probe, g_shift);
    return ADM_OK;
}

extern "C" int adm_probe_shift_fwd(int B, int M, int P, const float* probe, const float* shift, float* out, void* stream) {
    if (B < 1 || M < 1 || !probe || !shift || !out) return ADM_ERR_ARG;
    if (P == 72) return launch_shift_fwd<Cfg72>(B, M, probe, shift, out, stream);
    if (P == 64) return launch_shift_fwd<Cfg64>(B, M, probe, shift, out, stream);
    return ADM_ERR_UNSUPPORTED;
}

extern "C" int adm_probe_shift_bwd(int B, int M, int P, const float* probe, const float* shift, const float* g_out,
                                   float* g_probe, float* g_shift, void* stream) {
    if (B < 1 || M < 1 || !probe || !shift || !g_out || !g_shift) return ADM_ERR_ARG;
    if (P == 72) return launch_shift_bwd<Cfg72>(B, M, probe, shift, g_out, g_probe, g_shift, stream);
    if (P == 64) return launch_shift_bwd<Cfg64>(B, M, probe, shift, g_out, g_probe, g_shift, stream);
    return ADM_ERR_UNSUPPORTED;
}
