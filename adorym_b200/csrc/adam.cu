// Fused Adam update and L1 / TV regulariser kernels (HBM-bound elementwise work).
// Reference: adorym/optimizers.py:309-318 (Adam; eps outside the sqrt, bias correction with i_batch + 1),
// ptychography.py:1138-1142 (non-negativity / phase-only / absorption-only), array_ops.py:239-244 (mask),
// regularizers.py:30-46,64-83,95-110 + util.py:1427-1440 (L1, reweighted L1, TV with periodic roll).
#include "fft_passes.cuh"
#include "../../include/adorym_b200.h"

struct AdamArgs {
    float step_size, b1, b2, omb1, omb2, inv_q1, inv_q2, eps;
    int flags;
};

__global__ void __launch_bounds__(256) k_adam(float* x, float* m, float* v, const float* g, long long n2, AdamArgs a,
                                              const float* mask) {
    // one thread per (channel 0, channel 1) pair: n2 = n / 2 pairs (n odd handled by the host wrapper)
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) {
        float mk_ = (mask != nullptr) ? mask[i] : 1.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            long long j = 2 * i + c;
            float gg = g[j];
            float mm = a.b1 * m[j];
            mm = mm + a.omb1 * gg;
            float vv = a.b2 * v[j];
            vv = vv + a.omb2 * (gg * gg);
            float mhat = mm * a.inv_q1;
            float vhat = vv * a.inv_q2;
            float d = a.step_size * mhat / (sqrtf(vhat) + a.eps);
            float xx = x[j] - d;
            if (a.flags & 1) xx = fmaxf(xx, 0.f);
            if ((a.flags & 2) && c == 0) xx = 0.f;
            if ((a.flags & 4) && c == 1) xx = 0.f;
            if (mask != nullptr) xx *= mk_;
            x[j] = xx; m[j] = mm; v[j] = vv;
        }
    }
}

__global__ void __launch_bounds__(256) k_adam_tail(float* x, float* m, float* v, const float* g, long long j, AdamArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float gg = g[j];
    float mm = a.b1 * m[j] + a.omb1 * gg;
    float vv = a.b2 * v[j] + a.omb2 * (gg * gg);
    float d = a.step_size * (mm * a.inv_q1) / (sqrtf(vv * a.inv_q2) + a.eps);
    float xx = x[j] - d;
    if (a.flags & 1) xx = fmaxf(xx, 0.f);
    x[j] = xx; m[j] = mm; v[j] = vv;
}

extern "C" int adm_adam_step(float* x, float* m, float* v, const float* g, long long n, float step_size, float b1, float b2,
                             float eps, int step, int flags, const float* mask, void* stream) {
    if (!x || !m || !v || !g || n < 1 || step < 0) return ADM_ERR_ARG;
    if (((flags & 6) || mask) && (n & 1)) return ADM_ERR_ARG;
    AdamArgs a;
    a.step_size = step_size; a.b1 = b1; a.b2 = b2; a.eps = eps; a.flags = flags;
    a.omb1 = (float)(1.0 - (double)b1);
    a.omb2 = (float)(1.0 - (double)b2);
    a.inv_q1 = (float)(1.0 / (1.0 - pow((double)b1, (double)(step + 1))));
    a.inv_q2 = (float)(1.0 / (1.0 - pow((double)b2, (double)(step + 1))));
    long long n2 = n / 2;
    if (n2 > 0) {
        long long blocks = (n2 + 255) / 256;
        if (blocks > 148LL * 32) blocks = 148LL * 32;   // grid-stride; a multiple of the SM count
        ADM_LAUNCH(k_adam, (unsigned)blocks, 256, 0, stream, x, m, v, g, n2, a, mask);
    }
    if (n & 1) ADM_LAUNCH(k_adam_tail, 1, 32, 0, stream, x, m, v, g, n - 1, a);
    return ADM_OK;
}

// ---- regulariser -------------------------------------------------------------------------------
ADM_DEV float sgnf(float a) { return (a > 0.f) ? 1.f : ((a < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(256) k_reg_l1tv(int Y, int X, int Z, const cpx* obj, const cpx* wl1, float ad, float ab,
                                                  float gm, float* value, cpx* g_obj) {
#ifndef ADM_HOST_EMU
    __shared__ float red[256];
#else
    float* red = (float*)emu_smem;
#endif
    const long long N = (long long)Y * X * Z;
    const float invN = 1.0f / (float)N;
    float acc = 0.f;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < N; i += stride) {
        int z = (int)(i % Z);
        int x = (int)((i / Z) % X);
        int y = (int)(i / ((long long)Z * X));
        cpx c = obj[i];
        float w0 = 1.f, w1 = 1.f;
        if (wl1 != nullptr) { cpx w = wl1[i]; w0 = w.x; w1 = w.y; }
        float gd = ad * w0 * sgnf(c.x) * invN;
        float gb = ab * w1 * sgnf(c.y) * invN;
        acc += (ad * w0 * fabsf(c.x) + ab * w1 * fabsf(c.y)) * invN;
        if (gm != 0.f) {
            // TV: sum over the three axes of |roll(a, 1) - a| / N (periodic), per channel
            long long im, ip;
            cpx a, b;
            im = (long long)((y + Y - 1) % Y) * X * Z + (long long)x * Z + z;
            ip = (long long)((y + 1) % Y) * X * Z + (long long)x * Z + z;
            a = obj[im]; b = obj[ip];
            acc += gm * (fabsf(a.x - c.x) + fabsf(a.y - c.y)) * invN;
            gd += gm * (sgnf(c.x - a.x) + sgnf(c.x - b.x)) * invN;
            gb += gm * (sgnf(c.y - a.y) + sgnf(c.y - b.y)) * invN;
            im = (long long)y * X * Z + (long long)((x + X - 1) % X) * Z + z;
            ip = (long long)y * X * Z + (long long)((x + 1) % X) * Z + z;
            a = obj[im]; b = obj[ip];
            acc += gm * (fabsf(a.x - c.x) + fabsf(a.y - c.y)) * invN;
            gd += gm * (sgnf(c.x - a.x) + sgnf(c.x - b.x)) * invN;
            gb += gm * (sgnf(c.y - a.y) + sgnf(c.y - b.y)) * invN;
            im = (long long)y * X * Z + (long long)x * Z + (z + Z - 1) % Z;
            ip = (long long)y * X * Z + (long long)x * Z + (z + 1) % Z;
            a = obj[im]; b = obj[ip];
            acc += gm * (fabsf(a.x - c.x) + fabsf(a.y - c.y)) * invN;
            gd += gm * (sgnf(c.x - a.x) + sgnf(c.x - b.x)) * invN;
            gb += gm * (sgnf(c.y - a.y) + sgnf(c.y - b.y)) * invN;
        }
        if (g_obj != nullptr) { cpx o = g_obj[i]; o.x += gd; o.y += gb; g_obj[i] = o; }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0 && value != nullptr) atomicAdd(value, red[0]);
}

extern "C" int adm_regularizer_l1tv(int Y, int X, int Z, const float* obj, const float* weight_l1, float alpha_d, float alpha_b,
                                    float gamma, float* value, float* g_obj, void* stream) {
    if (Y < 1 || X < 1 || Z < 1 || !obj) return ADM_ERR_ARG;
    long long N = (long long)Y * X * Z;
    long long blocks = (N + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    ADM_LAUNCH(k_reg_l1tv, (unsigned)blocks, 256, 256 * sizeof(float), stream, Y, X, Z, (const cpx*)obj, (const cpx*)weight_l1,
               alpha_d, alpha_b, gamma, value, (cpx*)g_obj);
    return ADM_OK;
}
