// Affine resampling of a stack of 2-D planes, its transpose and its derivative with respect to the 2x3 matrices.
//
// Reference behaviour reproduced (PyTorch backend of the reference):
//   adorym/wrappers.py:1158-1174  affine_transform(arr[N,H,W], transform[2,3])  -- refinable affine alignment of the
//                                 measured holograms (forward_model.py:1066-1072, optimize_prj_affine)
//   adorym/wrappers.py:1177-1217  rotate(arr[Y,X,Z,C], theta, axis)             -- tilt refinement with d/d theta
//                                 (forward_model.py:274-276, optimize_tilt)
// both of which are F.affine_grid(m, size, align_corners=False) followed by F.grid_sample(mode='bilinear',
// padding_mode='border', align_corners=False).  The reference differentiates them with autograd; here the transpose
// (gradient with respect to the resampled array) and the matrix gradient are explicit kernels.
//
// The array is a dense [D0][D1][D2] array of C-vectors (C = 1: frames, C = 2: the (delta, beta) / (re, im) object); which
// memory dimension plays the role of the plane index N, the rows H and the columns W is a parameter, so the three
// rotations of a tilt (about y, x and z of the [Y][X][Z] object) need no permuted copies.
#include "adm_common.cuh"
#include "fft_passes.cuh"
#include "../../include/adorym_b200.h"

#define AFF_NT 256

struct AffGeom {
    int N, H, W;
    long long sn, sh, sw;      // strides in C-vectors
    int w_fast;                // 1: consecutive threads walk w, 0: they walk h
};

// base grid of affine_grid(align_corners=False): torch.linspace(-1, 1, n) * (n - 1) / n in fp32
ADM_DEV float aff_base(int i, int n) {
    if (n <= 1) return 0.f;
    const float step = 2.f / (float)(n - 1);
    const float lin = (i < n / 2) ? (-1.f + step * (float)i) : (1.f - step * (float)(n - 1 - i));
    return lin * (float)(n - 1) / (float)n;
}

struct AffTap {
    long long o00, o01, o10, o11;   // offsets of the four neighbours (clamped to valid memory)
    float w00, w01, w10, w11;       // bilinear weights (0 where the neighbour lies outside the plane)
    float ax0, ax1, ay0, ay1;       // (x0 + 1 - ix), (ix - x0), (y0 + 1 - iy), (iy - y0)
    float mx, my;                   // d ix / d x', d iy / d y' (0 where the border clamp is active)
    bool v01, v10, v11;
};

ADM_DEV AffTap aff_taps(const AffGeom& g, const float* m, float bx, float by) {
    const float xs = fmaf(by, m[1], bx * m[0]) + m[2];
    const float ys = fmaf(by, m[4], bx * m[3]) + m[5];
    float ix = ((xs + 1.f) * (float)g.W - 1.f) / 2.f;
    float iy = ((ys + 1.f) * (float)g.H - 1.f) / 2.f;
    AffTap t;
    // clip_coordinates_set_grad (ATen GridSampler.h): gradient 0 at and beyond the border
    t.mx = (ix <= 0.f || ix >= (float)(g.W - 1)) ? 0.f : 0.5f * (float)g.W;
    t.my = (iy <= 0.f || iy >= (float)(g.H - 1)) ? 0.f : 0.5f * (float)g.H;
    ix = fminf((float)(g.W - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(g.H - 1), fmaxf(iy, 0.f));
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const bool vx1 = x0 + 1 <= g.W - 1, vy1 = y0 + 1 <= g.H - 1;
    const int x1 = vx1 ? x0 + 1 : x0, y1 = vy1 ? y0 + 1 : y0;
    t.ax0 = (x0f + 1.f) - ix; t.ax1 = ix - x0f;
    t.ay0 = (y0f + 1.f) - iy; t.ay1 = iy - y0f;
    t.o00 = y0 * g.sh + x0 * g.sw; t.o01 = y0 * g.sh + x1 * g.sw;
    t.o10 = y1 * g.sh + x0 * g.sw; t.o11 = y1 * g.sh + x1 * g.sw;
    t.v01 = vx1; t.v10 = vy1; t.v11 = vx1 && vy1;
    t.w00 = t.ax0 * t.ay0;
    t.w01 = vx1 ? t.ax1 * t.ay0 : 0.f;
    t.w10 = vy1 ? t.ax0 * t.ay1 : 0.f;
    t.w11 = t.v11 ? t.ax1 * t.ay1 : 0.f;
    return t;
}

ADM_DEV bool aff_index(const AffGeom& g, int& n, int& h, int& w) {
    const long long j = (long long)blockIdx.x * AFF_NT + threadIdx.x;
    n = blockIdx.y;
    if (j >= (long long)g.H * g.W) return false;
    if (g.w_fast) { h = (int)(j / g.W); w = (int)(j % g.W); }
    else { w = (int)(j / g.H); h = (int)(j % g.H); }
    return true;
}

template <int C> __global__ void __launch_bounds__(AFF_NT) k_affine_fwd(AffGeom g, const float* in, const float* mats,
                                                                        int per_plane, float* out) {
    int n, h, w;
    if (!aff_index(g, n, h, w)) return;
    const float* m = mats + (per_plane ? 6 * n : 0);
    const AffTap t = aff_taps(g, m, aff_base(w, g.W), aff_base(h, g.H));
    const float* pl = in + (size_t)n * g.sn * C;
    float* o = out + ((size_t)n * g.sn + (size_t)h * g.sh + (size_t)w * g.sw) * C;
#pragma unroll
    for (int c = 0; c < C; ++c)
        o[c] = pl[t.o00 * C + c] * t.w00 + pl[t.o01 * C + c] * t.w01 + pl[t.o10 * C + c] * t.w10 + pl[t.o11 * C + c] * t.w11;
}

template <int C> __global__ void __launch_bounds__(AFF_NT) k_affine_adj(AffGeom g, const float* g_out, const float* mats,
                                                                        int per_plane, float* g_in) {
    int n, h, w;
    if (!aff_index(g, n, h, w)) return;
    const float* m = mats + (per_plane ? 6 * n : 0);
    const AffTap t = aff_taps(g, m, aff_base(w, g.W), aff_base(h, g.H));
    float* pl = g_in + (size_t)n * g.sn * C;
    const float* go = g_out + ((size_t)n * g.sn + (size_t)h * g.sh + (size_t)w * g.sw) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float v = go[c];
        if (t.w00 != 0.f) atomicAdd(pl + t.o00 * C + c, v * t.w00);
        if (t.w01 != 0.f) atomicAdd(pl + t.o01 * C + c, v * t.w01);
        if (t.w10 != 0.f) atomicAdd(pl + t.o10 * C + c, v * t.w10);
        if (t.w11 != 0.f) atomicAdd(pl + t.o11 * C + c, v * t.w11);
    }
}

// g_mats[n][6] += sum over the plane of g_out . d out / d m   (grid_sampler_2d_backward's grad_grid chained through the
// linear affine_grid: d x' / d (m0, m1, m2) = (bx, by, 1), d y' / d (m3, m4, m5) = (bx, by, 1))
template <int C> __global__ void __launch_bounds__(AFF_NT) k_affine_dmat(AffGeom g, const float* in, const float* g_out,
                                                                         const float* mats, int per_plane, float* g_mats) {
    ADM_SMEM_DECL
    float (*red)[AFF_NT] = (float (*)[AFF_NT])smem_raw;       // [6][AFF_NT]
    int n, h, w;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (aff_index(g, n, h, w)) {
        const float* m = mats + (per_plane ? 6 * n : 0);
        const float bx = aff_base(w, g.W), by = aff_base(h, g.H);
        const AffTap t = aff_taps(g, m, bx, by);
        const float* pl = in + (size_t)n * g.sn * C;
        const float* go = g_out + ((size_t)n * g.sn + (size_t)h * g.sh + (size_t)w * g.sw) * C;
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float v00 = pl[t.o00 * C + c];
            const float v01 = t.v01 ? pl[t.o01 * C + c] : 0.f;
            const float v10 = t.v10 ? pl[t.o10 * C + c] : 0.f;
            const float v11 = t.v11 ? pl[t.o11 * C + c] : 0.f;
            const float gv = go[c];
            gx += gv * ((v01 - v00) * t.ay0 + (v11 - v10) * t.ay1);
            gy += gv * ((v10 - v00) * t.ax0 + (v11 - v01) * t.ax1);
        }
        gx *= t.mx; gy *= t.my;
        acc[0] = gx * bx; acc[1] = gx * by; acc[2] = gx;
        acc[3] = gy * bx; acc[4] = gy * by; acc[5] = gy;
    }
    n = blockIdx.y;
#pragma unroll
    for (int k = 0; k < 6; ++k) red[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int s = AFF_NT / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
#pragma unroll
            for (int k = 0; k < 6; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x < 6) atomicAdd(g_mats + 6 * n + threadIdx.x, red[threadIdx.x][0]);
}

static int aff_geom(int D0, int D1, int D2, int C, int dn, int dh, int dw, AffGeom& g, dim3& grid) {
    if (D0 < 1 || D1 < 1 || D2 < 1 || (C != 1 && C != 2)) return ADM_ERR_ARG;
    if (dn < 0 || dn > 2 || dh < 0 || dh > 2 || dw < 0 || dw > 2 || dn == dh || dn == dw || dh == dw) return ADM_ERR_ARG;
    const int dims[3] = {D0, D1, D2};
    const long long str[3] = {(long long)D1 * D2, (long long)D2, 1};
    g.N = dims[dn]; g.H = dims[dh]; g.W = dims[dw];
    g.sn = str[dn]; g.sh = str[dh]; g.sw = str[dw];
    g.w_fast = g.sw < g.sh ? 1 : 0;
    if (g.N > 65535) return ADM_ERR_UNSUPPORTED;
    grid = dim3((unsigned)(((long long)g.H * g.W + AFF_NT - 1) / AFF_NT), (unsigned)g.N, 1);
    return ADM_OK;
}

extern "C" int adm_affine_resample_fwd(int D0, int D1, int D2, int C, int dim_n, int dim_h, int dim_w, const float* in,
                                       const float* mats, int mat_per_plane, float* out, void* stream) {
    AffGeom g; dim3 grid;
    if (!in || !mats || !out) return ADM_ERR_ARG;
    int rc = aff_geom(D0, D1, D2, C, dim_n, dim_h, dim_w, g, grid);
    if (rc) return rc;
    if (C == 1) ADM_LAUNCH(k_affine_fwd<1>, grid, AFF_NT, 0, stream, g, in, mats, mat_per_plane, out);
    else ADM_LAUNCH(k_affine_fwd<2>, grid, AFF_NT, 0, stream, g, in, mats, mat_per_plane, out);
    return ADM_OK;
}

extern "C" int adm_affine_resample_adj(int D0, int D1, int D2, int C, int dim_n, int dim_h, int dim_w, const float* g_out,
                                       const float* mats, int mat_per_plane, float* g_in, void* stream) {
    AffGeom g; dim3 grid;
    if (!g_out || !mats || !g_in) return ADM_ERR_ARG;
    int rc = aff_geom(D0, D1, D2, C, dim_n, dim_h, dim_w, g, grid);
    if (rc) return rc;
    if (C == 1) ADM_LAUNCH(k_affine_adj<1>, grid, AFF_NT, 0, stream, g, g_out, mats, mat_per_plane, g_in);
    else ADM_LAUNCH(k_affine_adj<2>, grid, AFF_NT, 0, stream, g, g_out, mats, mat_per_plane, g_in);
    return ADM_OK;
}

extern "C" int adm_affine_resample_dmat(int D0, int D1, int D2, int C, int dim_n, int dim_h, int dim_w, const float* in,
                                        const float* g_out, const float* mats, int mat_per_plane, float* g_mats,
                                        void* stream) {
    AffGeom g; dim3 grid;
    if (!in || !g_out || !mats || !g_mats) return ADM_ERR_ARG;
    int rc = aff_geom(D0, D1, D2, C, dim_n, dim_h, dim_w, g, grid);
    if (rc) return rc;
    if (C == 1) ADM_LAUNCH(k_affine_dmat<1>, grid, AFF_NT, 6 * AFF_NT * sizeof(float), stream, g, in, g_out, mats, mat_per_plane, g_mats);
    else ADM_LAUNCH(k_affine_dmat<2>, grid, AFF_NT, 6 * AFF_NT * sizeof(float), stream, g, in, g_out, mats, mat_per_plane, g_mats);
    return ADM_OK;
}
