// Tomographic rotation about the y axis and its exact transpose.
// Reference: adorym/util.py:536-552 apply_rotation -> wrappers.py:1105-1147 grid_sample ->
// torch grid_sample(mode='bilinear', padding_mode='border', align_corners=False); the transpose is what
// autograd's grid_sampler_2d_backward computes (prototype in the reference: util.py:623-696).
// Input volume is the reference layout [Y][X][Z][2]; the rotated volume is written SLICE-MAJOR [Z][Y][X][2]
// so that the multislice kernels read contiguous P*8-byte rows per slice.
#include "fft_passes.cuh"
#include "../../include/adorym_b200.h"

struct Taps {
    int i00, i01, i10, i11;       // offsets (x*Z + z) of the four taps (clamped to valid memory)
    float w00, w01, w10, w11;     // weights, 0 for taps outside the plane
};

// ATen GridSampler: unnormalise (align_corners=False), clip to the border, floor, nw/ne/sw/se weights.
ADM_DEV Taps make_taps(float gz, float gx, int X, int Z) {
    float iz = ((gz + 1.f) * (float)Z - 1.f) / 2.f;
    float ix = ((gx + 1.f) * (float)X - 1.f) / 2.f;
    iz = fminf((float)(Z - 1), fmaxf(iz, 0.f));
    ix = fminf((float)(X - 1), fmaxf(ix, 0.f));
    float z0f = floorf(iz), x0f = floorf(ix);
    int z0 = (int)z0f, x0 = (int)x0f, z1 = z0 + 1, x1 = x0 + 1;
    float wz1 = iz - z0f, wz0 = (z0f + 1.f) - iz;
    float wx1 = ix - x0f, wx0 = (x0f + 1.f) - ix;
    bool vz1 = z1 <= Z - 1, vx1 = x1 <= X - 1;
    int z1c = vz1 ? z1 : z0, x1c = vx1 ? x1 : x0;
    Taps t;
    t.i00 = x0 * Z + z0;  t.w00 = wx0 * wz0;
    t.i01 = x0 * Z + z1c; t.w01 = vz1 ? wx0 * wz1 : 0.f;
    t.i10 = x1c * Z + z0; t.w10 = vx1 ? wx1 * wz0 : 0.f;
    t.i11 = x1c * Z + z1c; t.w11 = (vx1 && vz1) ? wx1 * wz1 : 0.f;
    return t;
}

#define ROT_YCHUNK 8

// Thread -> output voxel (x, z) of a compact 16 x 8 tile (each warp an 8 x 4 sub-tile): the taps of a warp then fall
// in a ~10 x 7 patch of the source plane (a few dozen 32-byte sectors shared between the lanes) instead of along a
// 32-voxel line (one sector per lane and tap); the stores are 64-byte runs along x.
#define ROT_TX 16
#define ROT_TZ 8
ADM_DEV void rot_xz(int& x, int& z) {
    const int t = threadIdx.x, w = t >> 5, l = t & 31;
    x = blockIdx.x * ROT_TX + ((w & 1) << 3) + (l & 7);
    z = blockIdx.y * ROT_TZ + ((w >> 1) << 2) + (l >> 3);
}

// grid (ceil(X/16), ceil(Z/8), ceil(Y/ROT_YCHUNK)); loop over a chunk of y with the taps held in registers
__global__ void __launch_bounds__(128) k_rotate_fwd(int Y, int X, int Z, const cpx* obj, const float* grid, cpx* rot) {
    int x, z;
    rot_xz(x, z);
    if (x >= X || z >= Z) return;
    const float* g = grid + 2 * ((size_t)x * Z + z);
    Taps t = make_taps(g[0], g[1], X, Z);
    const int y0 = blockIdx.z * ROT_YCHUNK;
    cpx a[ROT_YCHUNK], b[ROT_YCHUNK], c[ROT_YCHUNK], d[ROT_YCHUNK];
    // all taps of the y chunk are requested before the first is used (8 x 4 independent loads in flight per thread)
#pragma unroll
    for (int k = 0; k < ROT_YCHUNK; ++k) {
        const int y = (y0 + k < Y) ? y0 + k : Y - 1;
        const cpx* pl = obj + (size_t)y * X * Z;
        a[k] = pl[t.i00]; b[k] = pl[t.i01]; c[k] = pl[t.i10]; d[k] = pl[t.i11];
    }
#pragma unroll
    for (int k = 0; k < ROT_YCHUNK; ++k) {
        const int y = y0 + k;
        if (y < Y) {
            cpx r;
            r.x = a[k].x * t.w00 + b[k].x * t.w01 + c[k].x * t.w10 + d[k].x * t.w11;
            r.y = a[k].y * t.w00 + b[k].y * t.w01 + c[k].y * t.w10 + d[k].y * t.w11;
            rot[((size_t)z * Y + y) * X + x] = r;
        }
    }
}

ADM_DEV void red2(cpx* addr, float a, float b) {
#ifndef ADM_HOST_EMU
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a), "f"(b) : "memory");
#else
    atomicAdd(&addr->x, a);
    atomicAdd(&addr->y, b);
#endif
}

__global__ void __launch_bounds__(128) k_rotate_adj(int Y, int X, int Z, const cpx* g_rot, const float* grid, cpx* g_obj) {
    int x, z;
    rot_xz(x, z);
    if (x >= X || z >= Z) return;
    const float* g = grid + 2 * ((size_t)x * Z + z);
    Taps t = make_taps(g[0], g[1], X, Z);
    const int y0 = blockIdx.z * ROT_YCHUNK;
    cpx vv[ROT_YCHUNK];
#pragma unroll
    for (int k = 0; k < ROT_YCHUNK; ++k) {
        const int y = (y0 + k < Y) ? y0 + k : Y - 1;
        vv[k] = g_rot[((size_t)z * Y + y) * X + x];
    }
#pragma unroll
    for (int k = 0; k < ROT_YCHUNK; ++k) {
        const int y = y0 + k;
        if (y >= Y) break;
        const cpx v = vv[k];
        cpx* pl = g_obj + (size_t)y * X * Z;
        if (t.w00 != 0.f) red2(pl + t.i00, v.x * t.w00, v.y * t.w00);
        if (t.w01 != 0.f) red2(pl + t.i01, v.x * t.w01, v.y * t.w01);
        if (t.w10 != 0.f) red2(pl + t.i10, v.x * t.w10, v.y * t.w10);
        if (t.w11 != 0.f) red2(pl + t.i11, v.x * t.w11, v.y * t.w11);
    }
}

extern "C" int adm_rotate_fwd(int Y, int X, int Z, const float* obj, const float* grid, float* rot, void* stream) {
    if (Y < 1 || X < 1 || Z < 1 || !obj || !grid || !rot) return ADM_ERR_ARG;
    if ((Z + ROT_TZ - 1) / ROT_TZ > 65535 || (Y + ROT_YCHUNK - 1) / ROT_YCHUNK > 65535) return ADM_ERR_UNSUPPORTED;
    dim3 gr((X + ROT_TX - 1) / ROT_TX, (Z + ROT_TZ - 1) / ROT_TZ, (Y + ROT_YCHUNK - 1) / ROT_YCHUNK);
    ADM_LAUNCH(k_rotate_fwd, gr, 128, 0, stream, Y, X, Z, (const cpx*)obj, grid, (cpx*)rot);
    return ADM_OK;
}

extern "C" int adm_rotate_adj(int Y, int X, int Z, const float* g_rot, const float* grid, float* g_obj, void* stream) {
    if (Y < 1 || X < 1 || Z < 1 || !g_rot || !grid || !g_obj) return ADM_ERR_ARG;
    if ((Z + ROT_TZ - 1) / ROT_TZ > 65535 || (Y + ROT_YCHUNK - 1) / ROT_YCHUNK > 65535) return ADM_ERR_UNSUPPORTED;
    dim3 gr((X + ROT_TX - 1) / ROT_TX, (Z + ROT_TZ - 1) / ROT_TZ, (Y + ROT_YCHUNK - 1) / ROT_YCHUNK);
    ADM_LAUNCH(k_rotate_adj, gr, 128, 0, stream, Y, X, Z, (const cpx*)g_rot, grid, (cpx*)g_obj);
    return ADM_OK;
}
