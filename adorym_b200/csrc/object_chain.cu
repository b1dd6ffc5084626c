// Object -> transmission volume in ONE pass, and its transpose: the per-step chain
//     obj [Y][X][Z][2]  --rotate about y (bilinear, border)-->  rot [Z][Y][X][2]  --T' = T(rot) - 1-->  tvol [Z][Y][Xp][2]
// without the intermediate rotated volume (rotate.cu + transmission.cu write and re-read 134 MB at 256^3 in each
// direction).  Reference: adorym/util.py:536-552 apply_rotation -> wrappers.py:1105-1147 grid_sample (bilinear,
// border, align_corners=False) followed, inside multislice_propagate_batch, by propagate.py:241 / :246 (transmission
// of every slice); the transpose is what torch.autograd.grad computes through both (wrappers.py:300-331).
//   forward : out voxel (z, y, x) = T'( sum_taps w obj[y][x'][z'] )                       (grid == NULL: identity taps,
//             i.e. a pure [Y][X][Z] -> [Z][Y][X] transposition: the rotate_out_of_loop schedule, ptychography.py:1011)
//   adjoint : g_rot(z, y, x) = chain rule of g_tvol through T (transmission.cu), scattered to the 4 taps of g_obj with
//             red.global.add.v2.f32; g_obj is NOT cleared here: the caller initialises it with the regulariser gradient
//             (adm_regularizer_l1tv_range, accumulate = 0) or with zeros.  Both run on a y range [y0, y1) so that the
//             gradient all-reduce can start on finished y slabs (the rotation is about y: planes are independent).
#include "fft_passes.cuh"
#include "../../include/adorym_b200.h"

struct OTaps {
    int i00, i01, i10, i11;       // offsets (x*Z + z) of the four taps (clamped to valid memory)
    float w00, w01, w10, w11;     // weights, 0 for taps outside the plane
};

// ATen GridSampler: unnormalise (align_corners=False), clip to the border, floor, nw/ne/sw/se weights (as rotate.cu).
ADM_DEV OTaps oc_make_taps(const float* grid, int x, int z, int X, int Z) {
    OTaps t;
    if (grid == nullptr) {
        t.i00 = t.i01 = t.i10 = t.i11 = x * Z + z;
        t.w00 = 1.f; t.w01 = t.w10 = t.w11 = 0.f;
        return t;
    }
    const float* g = grid + 2 * ((size_t)x * Z + z);
    float iz = ((g[0] + 1.f) * (float)Z - 1.f) / 2.f;
    float ix = ((g[1] + 1.f) * (float)X - 1.f) / 2.f;
    iz = fminf((float)(Z - 1), fmaxf(iz, 0.f));
    ix = fminf((float)(X - 1), fmaxf(ix, 0.f));
    float z0f = floorf(iz), x0f = floorf(ix);
    int z0 = (int)z0f, x0 = (int)x0f, z1 = z0 + 1, x1 = x0 + 1;
    float wz1 = iz - z0f, wz0 = (z0f + 1.f) - iz;
    float wx1 = ix - x0f, wx0 = (x0f + 1.f) - ix;
    bool vz1 = z1 <= Z - 1, vx1 = x1 <= X - 1;
    int z1c = vz1 ? z1 : z0, x1c = vx1 ? x1 : x0;
    t.i00 = x0 * Z + z0;  t.w00 = wx0 * wz0;
    t.i01 = x0 * Z + z1c; t.w01 = vz1 ? wx0 * wz1 : 0.f;
    t.i10 = x1c * Z + z0; t.w10 = vx1 ? wx1 * wz0 : 0.f;
    t.i11 = x1c * Z + z1c; t.w11 = (vx1 && vz1) ? wx1 * wz1 : 0.f;
    return t;
}

// T - 1 without cancellation (transmission.cu)
ADM_DEV cpx oc_transmission_m1(cpx o, int unknown_type, float k1, float sgn) {
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

#define OC_YCHUNK 8
#define OC_TX 16
#define OC_TZ 8
// thread -> output voxel (x, z) of a compact 16 x 8 tile (each warp an 8 x 4 sub-tile), see rotate.cu
ADM_DEV void oc_xz(int& x, int& z) {
    const int t = threadIdx.x, w = t >> 5, l = t & 31;
    x = blockIdx.x * OC_TX + ((w & 1) << 3) + (l & 7);
    z = blockIdx.y * OC_TZ + ((w >> 1) << 2) + (l >> 3);
}

// `bin` > 1 (propagate.py:207-254, `binning`): output slice zb is the transmission of the SUM of (delta, beta) over the
// object slices zb*bin .. zb*bin + bin - 1 (delta_beta objects); z counts OUTPUT slices, Zo = ceil(Z / bin).
template <bool BIN1> __global__ void __launch_bounds__(128) k_object_to_transmission(int Y, int X, int Z, int Xp, const cpx* obj,
                                                                                     const float* grid, int unknown_type, float k1,
                                                                                     float sgn, cpx* tvol, int bin_) {
    const int bin = BIN1 ? 1 : bin_;
    int x, zb;
    oc_xz(x, zb);
    const int Zo = (Z + bin - 1) / bin;
    if (x >= X || zb >= Zo) return;
    const int y0 = blockIdx.z * OC_YCHUNK;
    const bool one_tap = (grid == nullptr);
    cpx acc[OC_YCHUNK];
#pragma unroll
    for (int k = 0; k < OC_YCHUNK; ++k) acc[k] = mk(0.f, 0.f);
    const int z_end = (zb * bin + bin < Z) ? zb * bin + bin : Z;
    for (int z = zb * bin; z < z_end; ++z) {
        const OTaps t = oc_make_taps(grid, x, z, X, Z);
        cpx a[OC_YCHUNK], b[OC_YCHUNK], c[OC_YCHUNK], d[OC_YCHUNK];
#pragma unroll
        for (int k = 0; k < OC_YCHUNK; ++k) {
            const int y = (y0 + k < Y) ? y0 + k : Y - 1;
            const cpx* pl = obj + (size_t)y * X * Z;
            a[k] = pl[t.i00];
            if (!one_tap) { b[k] = pl[t.i01]; c[k] = pl[t.i10]; d[k] = pl[t.i11]; }
        }
#pragma unroll
        for (int k = 0; k < OC_YCHUNK; ++k) {
            cpx r = a[k];
            if (!one_tap) {
                r.x = a[k].x * t.w00 + b[k].x * t.w01 + c[k].x * t.w10 + d[k].x * t.w11;
                r.y = a[k].y * t.w00 + b[k].y * t.w01 + c[k].y * t.w10 + d[k].y * t.w11;
            }
            if (BIN1) acc[k] = r;
            else { acc[k].x += r.x; acc[k].y += r.y; }
        }
    }
#pragma unroll
    for (int k = 0; k < OC_YCHUNK; ++k) {
        const int y = y0 + k;
        if (y < Y) {
            cpx* row = tvol + ((size_t)zb * Y + y) * Xp;
            row[x] = oc_transmission_m1(acc[k], unknown_type, k1, sgn);
            if (x == X - 1 && Xp > X) row[X] = mk(0.f, 0.f);      // pad column of the TMA row pitch
        }
    }
}

ADM_DEV void oc_red2(cpx* addr, float a, float b) {
#ifndef ADM_HOST_EMU
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a), "f"(b) : "memory");
#else
    atomicAdd(&addr->x, a);
    atomicAdd(&addr->y, b);
#endif
}

// g_tvol: margined gradient volume [Zo][Y + mg][Xg][2]; grid.z covers the y range [y0, y1) in chunks of OC_YCHUNK.
// bin > 1: every object slice of the bin receives the gradient of the summed (delta, beta).
__global__ void __launch_bounds__(128) k_transmission_to_object_adj(int Y, int X, int Z, int Xp, int mg, int Xg, const cpx* g_tvol,
                                                                    const cpx* tvol, const float* grid, int unknown_type,
                                                                    float k1, float sgn, cpx* g_obj, int y_begin, int y_end,
                                                                    int bin) {
    int x, zb;
    oc_xz(x, zb);
    const int Zo = (Z + bin - 1) / bin;
    if (x >= X || zb >= Zo) return;
    const int y0 = y_begin + blockIdx.z * OC_YCHUNK;
    cpx gv[OC_YCHUNK], tv[OC_YCHUNK];
#pragma unroll
    for (int k = 0; k < OC_YCHUNK; ++k) {
        const int y = (y0 + k < y_end) ? y0 + k : y_end - 1;
        gv[k] = g_tvol[((size_t)zb * (Y + mg) + y + mg) * Xg + x + mg];
        if (unknown_type == ADM_DELTA_BETA) tv[k] = tvol[((size_t)zb * Y + y) * Xp + x];
    }
#pragma unroll
    for (int k = 0; k < OC_YCHUNK; ++k) {
        if (unknown_type == ADM_DELTA_BETA) {
            // dL/ddelta = -s k1 Im(g_T conj(T)), dL/dbeta = -k1 Re(g_T conj(T))   (transmission.cu)
            cpx u = cmulc(gv[k], mk(tv[k].x + 1.f, tv[k].y));
            gv[k] = mk(-sgn * k1 * u.y, -k1 * u.x);
        }
    }
    const int z_end = (zb * bin + bin < Z) ? zb * bin + bin : Z;
    for (int z = zb * bin; z < z_end; ++z) {
        const OTaps t = oc_make_taps(grid, x, z, X, Z);
#pragma unroll
        for (int k = 0; k < OC_YCHUNK; ++k) {
            const int y = y0 + k;
            if (y >= y_end) break;
            const cpx v = gv[k];
            cpx* pl = g_obj + (size_t)y * X * Z;
            if (t.w00 != 0.f) oc_red2(pl + t.i00, v.x * t.w00, v.y * t.w00);
            if (t.w01 != 0.f) oc_red2(pl + t.i01, v.x * t.w01, v.y * t.w01);
            if (t.w10 != 0.f) oc_red2(pl + t.i10, v.x * t.w10, v.y * t.w10);
            if (t.w11 != 0.f) oc_red2(pl + t.i11, v.x * t.w11, v.y * t.w11);
        }
    }
}

static int oc_check(int Y, int X, int Z) {
    if (Y < 1 || X < 1 || Z < 1) return ADM_ERR_ARG;
    if ((Z + OC_TZ - 1) / OC_TZ > 65535 || (Y + OC_YCHUNK - 1) / OC_YCHUNK > 65535) return ADM_ERR_UNSUPPORTED;
    return ADM_OK;
}

extern "C" int adm_object_to_transmission(int Y, int X, int Z, const float* obj, const float* grid, int unknown_type, float k1,
                                          int sign_convention, int binning, float* tvol, void* stream) {
    if (!obj || !tvol || binning < 1) return ADM_ERR_ARG;
    if (unknown_type != ADM_DELTA_BETA && unknown_type != ADM_REAL_IMAG) return ADM_ERR_ARG;
    if (binning > 1 && unknown_type != ADM_DELTA_BETA) return ADM_ERR_UNSUPPORTED;
    int rc = oc_check(Y, X, Z);
    if (rc) return rc;
    const int Zo = (Z + binning - 1) / binning;
    dim3 gr((X + OC_TX - 1) / OC_TX, (Zo + OC_TZ - 1) / OC_TZ, (Y + OC_YCHUNK - 1) / OC_YCHUNK);
    if (binning == 1)
        ADM_LAUNCH(k_object_to_transmission<true>, gr, 128, 0, stream, Y, X, Z, ADM_TVOL_PITCH(X), (const cpx*)obj, grid,
                   unknown_type, k1, (float)sign_convention, (cpx*)tvol, 1);
    else
        ADM_LAUNCH(k_object_to_transmission<false>, gr, 128, 0, stream, Y, X, Z, ADM_TVOL_PITCH(X), (const cpx*)obj, grid,
                   unknown_type, k1, (float)sign_convention, (cpx*)tvol, binning);
    return ADM_OK;
}

extern "C" int adm_transmission_to_object_adj(int Y, int X, int Z, int P, const float* g_tvol, const float* tvol,
                                              const float* grid, int unknown_type, float k1, int sign_convention,
                                              int binning, float* g_obj, int y_begin, int y_end, void* stream) {
    if (!g_tvol || !g_obj || P < 0 || binning < 1) return ADM_ERR_ARG;      // P = 0: un-margined gradient volume
    if (binning > 1 && unknown_type != ADM_DELTA_BETA) return ADM_ERR_UNSUPPORTED;
    if (unknown_type != ADM_DELTA_BETA && unknown_type != ADM_REAL_IMAG) return ADM_ERR_ARG;
    if (unknown_type == ADM_DELTA_BETA && !tvol) return ADM_ERR_ARG;
    if (y_begin < 0 || y_end > Y || y_begin >= y_end) return ADM_ERR_ARG;
    int rc = oc_check(Y, X, Z);
    if (rc) return rc;
    const int Zo = (Z + binning - 1) / binning;
    dim3 gr((X + OC_TX - 1) / OC_TX, (Zo + OC_TZ - 1) / OC_TZ, (y_end - y_begin + OC_YCHUNK - 1) / OC_YCHUNK);
    ADM_LAUNCH(k_transmission_to_object_adj, gr, 128, 0, stream, Y, X, Z, ADM_TVOL_PITCH(X), ADM_GVOL_MARGIN(P),
               ADM_GVOL_PITCH(X, P), (const cpx*)g_tvol, (const cpx*)tvol, grid, unknown_type, k1, (float)sign_convention,
               (cpx*)g_obj, y_begin, y_end, binning);
    return ADM_OK;
}
