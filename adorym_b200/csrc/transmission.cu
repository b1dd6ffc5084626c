// Transmission volume T' = T - 1 of a slice-major object volume, and the chain rule back to the object.
// Reference: adorym/propagate.py:241 (delta_beta: T = exp(-k1 beta) (cos(s k1 delta) - i sin(s k1 delta)),
// i.e. w.exp_complex(-k1 beta, -s k1 delta)) and :246 (real_imag: T = re + i im).  The multislice kernels consume
// T' so that (i) no transcendental runs in the per-slice loop although every voxel is visited by up to 36
// overlapping probe patches, (ii) vacuum outside the object is the TMA zero fill for both object types
// (util.py:1335-1350: (0,0) for delta_beta, (1,0) for real_imag), (iii) psi (1 + T') loses no precision for weak objects.
// Layout: rot [rows][X][2] dense  ->  tvol [rows][Xp][2], Xp = X rounded up to even (TMA row pitch must be a
// multiple of 16 bytes); rows = Nz * Y.
#include "adm_common.cuh"
#include "../../include/adorym_b200.h"

#ifndef ADM_HOST_EMU
#define ADM_LAUNCH_T(kern, grid, block, stream, ...)                                             \
    do {                                                                                         \
        kern<<<grid, block, 0, (cudaStream_t)(stream)>>>(__VA_ARGS__);                           \
        if (cudaGetLastError() != cudaSuccess) return ADM_ERR_CUDA;                              \
    } while (0)
#else
// grid-stride kernels: the emulation runs them with a handful of OS threads instead of one per CUDA thread
#define ADM_LAUNCH_T(kern, grid, block, stream, ...) emu_launch(dim3(1), dim3(8), 0, [=] { kern(__VA_ARGS__); })
#endif

// T - 1 computed without cancellation: e^{-b} cos(a) - 1 = expm1(-b) cos(a) + (cos(a) - 1), cos(a) - 1 = -2 sin^2(a/2)
ADM_DEV cpx transmission_m1(cpx o, int unknown_type, float k1, float sgn) {
    if (unknown_type == ADM_DELTA_BETA) {
        float a = -sgn * k1 * o.x;
        float em1 = expm1f(-k1 * o.y);
        float s, c, sh, ch;
        adm_sincos(a, &s, &c);
        adm_sincos(0.5f * a, &sh, &ch);
        float cm1 = -2.f * sh * sh;
        return mk(fmaf(em1, c, cm1), fmaf(em1, s, s));
    }
    return mk(o.x - 1.f, o.y);
}

__global__ void __launch_bounds__(256) k_transmission_fwd(int unknown_type, float k1, float sgn, long long rows, int X, int Xp,
                                                          const cpx* rot, cpx* tvol) {
    long long n = rows * Xp;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        long long r = i / Xp;
        int x = (int)(i - r * Xp);
        tvol[i] = (x < X) ? transmission_m1(rot[r * X + x], unknown_type, k1, sgn) : mk(0.f, 0.f);
    }
}

// g_rot = chain rule of dL/dT through T(delta, beta):  dT/ddelta = -i s k1 T,  dT/dbeta = -k1 T
//   dL/ddelta = -s k1 Im(g_T conj(T)),  dL/dbeta = -k1 Re(g_T conj(T));   real_imag: identity
// g_tvol is the MARGINED gradient volume [Nz][Y + mg][Xg][2] (object voxel (y, x) at (y + mg, x + mg)).
__global__ void __launch_bounds__(256) k_transmission_adj(int unknown_type, float k1, float sgn, int Nz, int Y, int X, int Xp,
                                                          int mg, int Xg, const cpx* g_tvol, const cpx* tvol, cpx* g_rot) {
    long long n = (long long)Nz * Y * X;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        long long r = i / X;
        int x = (int)(i - r * X);
        long long z = r / Y;
        int y = (int)(r - z * Y);
        cpx g = g_tvol[(z * (Y + mg) + y + mg) * Xg + x + mg];
        if (unknown_type == ADM_DELTA_BETA) {
            cpx tp = tvol[r * Xp + x];
            cpx u = cmulc(g, mk(tp.x + 1.f, tp.y));
            g_rot[i] = mk(-sgn * k1 * u.y, -k1 * u.x);
        } else {
            g_rot[i] = g;
        }
    }
}

extern "C" int adm_transmission_fwd(int unknown_type, float k1, int sign_convention, long long rows, int X, const float* rot,
                                    float* tvol, void* stream) {
    if (rows < 1 || X < 1 || !rot || !tvol) return ADM_ERR_ARG;
    if (unknown_type != ADM_DELTA_BETA && unknown_type != ADM_REAL_IMAG) return ADM_ERR_ARG;
    int Xp = ADM_TVOL_PITCH(X);
    long long n = rows * Xp;
    long long blocks = (n + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    ADM_LAUNCH_T(k_transmission_fwd, (unsigned)blocks, 256, stream, unknown_type, k1, (float)sign_convention, rows, X, Xp,
                 (const cpx*)rot, (cpx*)tvol);
    return ADM_OK;
}

extern "C" int adm_transmission_adj(int unknown_type, float k1, int sign_convention, int Nz, int Y, int X, int P,
                                    const float* g_tvol, const float* tvol, float* g_rot, void* stream) {
    if (Nz < 1 || Y < 1 || X < 1 || P < 1 || !g_tvol || !tvol || !g_rot) return ADM_ERR_ARG;
    if (unknown_type != ADM_DELTA_BETA && unknown_type != ADM_REAL_IMAG) return ADM_ERR_ARG;
    int Xp = ADM_TVOL_PITCH(X);
    long long n = (long long)Nz * Y * X;
    long long blocks = (n + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    ADM_LAUNCH_T(k_transmission_adj, (unsigned)blocks, 256, stream, unknown_type, k1, (float)sign_convention, Nz, Y, X, Xp,
                 ADM_GVOL_MARGIN(P), ADM_GVOL_PITCH(X, P), (const cpx*)g_tvol, (const cpx*)tvol, (cpx*)g_rot);
    return ADM_OK;
}
