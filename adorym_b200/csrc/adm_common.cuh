// Common device helpers for the adorym_b200 kernels (sm_100a).
//
// The same sources also compile as plain C++ under -DADM_HOST_EMU (tests/emu/): one OS thread per
// CUDA thread, a pthread-style barrier for __syncthreads().  That build exists ONLY so that the
// kernel logic (index permutations, butterflies, adjoints) can be unit-tested in the GPU-less build
// container; it is never loaded by the product package.
#pragma once
#include <stdint.h>
#include <math.h>

#ifndef ADM_HOST_EMU
#include <cuda_runtime.h>
#define ADM_DEV __device__ __forceinline__
#define ADM_HD __host__ __device__ __forceinline__
#define ADM_HDM __host__ __device__ __forceinline__
#else
#include "emu_runtime.h"
#define ADM_DEV static inline
#define ADM_HD static inline
#define ADM_HDM inline
#endif

#define ADM_OK 0
#define ADM_ERR_ARG 1
#define ADM_ERR_UNSUPPORTED 2
#define ADM_ERR_CUDA 3

struct cpx {
    float x, y;
};

ADM_HD cpx mk(float a, float b) { cpx r; r.x = a; r.y = b; return r; }
ADM_HD cpx operator+(cpx a, cpx b) { return mk(a.x + b.x, a.y + b.y); }
ADM_HD cpx operator-(cpx a, cpx b) { return mk(a.x - b.x, a.y - b.y); }
ADM_HD cpx cmul(cpx a, cpx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
ADM_HD cpx cmulc(cpx a, cpx b) { return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
ADM_HD cpx cconj(cpx a) { return mk(a.x, -a.y); }
ADM_HD cpx cscale(cpx a, float s) { return mk(a.x * s, a.y * s); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> ADM_HD cpx mul_mi(cpx a) { return INV ? mk(-a.y, a.x) : mk(a.y, -a.x); }

#define ADM_SQRT1_2 0.70710678118654752440f
#define ADM_SQRT3_2 0.86602540378443864676f

// ---- radix-2/4/8 ---------------------------------------------------------------------------
// In-register DFT of 8 points, natural order in, natural order out.
// INV=false: X[k] = sum x[n] exp(-2 pi i n k / 8);  INV=true: conjugate twiddles (unnormalised).
template <bool INV> ADM_HD void dft4(cpx& d0, cpx& d1, cpx& d2, cpx& d3) {
    cpx e0 = d0 + d2, e1 = d0 - d2, e2 = d1 + d3, e3 = mul_mi<INV>(d1 - d3);
    d0 = e0 + e2; d2 = e0 - e2; d1 = e1 + e3; d3 = e1 - e3;
}

template <bool INV> ADM_HD void dft8(cpx* a) {
    cpx b0 = a[0] + a[4], c0 = a[0] - a[4];
    cpx b1 = a[1] + a[5], c1 = a[1] - a[5];
    cpx b2 = a[2] + a[6], c2 = a[2] - a[6];
    cpx b3 = a[3] + a[7], c3 = a[3] - a[7];
    dft4<INV>(b0, b1, b2, b3);   // even outputs X[0], X[2], X[4], X[6]
    a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
    // odd outputs: 4-point DFT of (c0, W8 c1, W8^2 c2, W8^3 c3), W8 = exp(-+ i pi/4).  With W8 c1 = s r1 and
    // W8^3 c3 = s r3 (s = 1/sqrt 2, r1 / r3 the unscaled rotations) the factor s is folded into the last butterfly as
    // FMAs, so the 8-point transform has no stand-alone multiplications.
    cpx r1, r3, e0, e1;
    if (!INV) {
        r1 = mk(c1.x + c1.y, c1.y - c1.x);
        r3 = mk(c3.y - c3.x, -(c3.x + c3.y));
        e0 = mk(c0.x + c2.y, c0.y - c2.x);      // c0 + (-i) c2
        e1 = mk(c0.x - c2.y, c0.y + c2.x);      // c0 - (-i) c2
    } else {
        r1 = mk(c1.x - c1.y, c1.x + c1.y);
        r3 = mk(-(c3.x + c3.y), c3.x - c3.y);
        e0 = mk(c0.x - c2.y, c0.y + c2.x);
        e1 = mk(c0.x + c2.y, c0.y - c2.x);
    }
    const cpx e2 = r1 + r3, t = r1 - r3;
    a[1] = mk(fmaf(ADM_SQRT1_2, e2.x, e0.x), fmaf(ADM_SQRT1_2, e2.y, e0.y));
    a[5] = mk(fmaf(-ADM_SQRT1_2, e2.x, e0.x), fmaf(-ADM_SQRT1_2, e2.y, e0.y));
    if (!INV) {   // e1 +- (-i) s t
        a[3] = mk(fmaf(ADM_SQRT1_2, t.y, e1.x), fmaf(-ADM_SQRT1_2, t.x, e1.y));
        a[7] = mk(fmaf(-ADM_SQRT1_2, t.y, e1.x), fmaf(ADM_SQRT1_2, t.x, e1.y));
    } else {      // e1 +- (+i) s t
        a[3] = mk(fmaf(-ADM_SQRT1_2, t.y, e1.x), fmaf(ADM_SQRT1_2, t.x, e1.y));
        a[7] = mk(fmaf(ADM_SQRT1_2, t.y, e1.x), fmaf(-ADM_SQRT1_2, t.x, e1.y));
    }
}

// ---- radix-3/9 -----------------------------------------------------------------------------
template <bool INV> ADM_HD void dft3(cpx& a0, cpx& a1, cpx& a2) {
    cpx t = a1 + a2;
    cpx u = mk(fmaf(-0.5f, t.x, a0.x), fmaf(-0.5f, t.y, a0.y));
    cpx d = a1 - a2;
    a0 = a0 + t;
    // forward: y1 = u - i c d, y2 = u + i c d (c = sqrt(3)/2, folded into the FMAs); inverse swaps
    cpx p = mk(fmaf(ADM_SQRT3_2, d.y, u.x), fmaf(-ADM_SQRT3_2, d.x, u.y));
    cpx q = mk(fmaf(-ADM_SQRT3_2, d.y, u.x), fmaf(ADM_SQRT3_2, d.x, u.y));
    if (!INV) { a1 = p; a2 = q; } else { a1 = q; a2 = p; }
}

// W9^j, j = 1, 2, 4 (cos, sin of 2 pi j / 9)
#define ADM_C9_1 0.76604444311897803520f
#define ADM_S9_1 0.64278760968653932632f
#define ADM_C9_2 0.17364817766693034885f
#define ADM_S9_2 0.98480775301220805937f
#define ADM_C9_4 (-0.93969262078590838405f)
#define ADM_S9_4 0.34202014332566873304f

template <bool INV> ADM_HD cpx tw9(cpx a, float c, float s) {
    // forward: multiply by exp(-i phi) = (c, -s); inverse: (c, +s)
    return INV ? mk(a.x * c - a.y * s, a.x * s + a.y * c) : mk(a.x * c + a.y * s, a.y * c - a.x * s);
}

// In-register DFT of 9 points, natural in / natural out.  n = 3 n1 + n2, k = k1 + 3 k2.
template <bool INV> ADM_HD void dft9(cpx* a) {
    // step 1: for each n2: radix-3 over n1 (inputs a[n2], a[3+n2], a[6+n2]) -> Y[k1][n2] in place
    dft3<INV>(a[0], a[3], a[6]);
    dft3<INV>(a[1], a[4], a[7]);
    dft3<INV>(a[2], a[5], a[8]);
    // twiddle W9^{n2 k1}: (k1=1,n2=1)->1 at a[4]; (k1=1,n2=2)->2 at a[5]; (k1=2,n2=1)->2 at a[7]; (k1=2,n2=2)->4 at a[8]
    a[4] = tw9<INV>(a[4], ADM_C9_1, ADM_S9_1);
    a[5] = tw9<INV>(a[5], ADM_C9_2, ADM_S9_2);
    a[7] = tw9<INV>(a[7], ADM_C9_2, ADM_S9_2);
    a[8] = tw9<INV>(a[8], ADM_C9_4, ADM_S9_4);
    // step 2: for each k1: radix-3 over n2 (inputs a[3k1 + n2]) -> X[k1 + 3 k2] at a[3 k1 + k2]
    dft3<INV>(a[0], a[1], a[2]);
    dft3<INV>(a[3], a[4], a[5]);
    dft3<INV>(a[6], a[7], a[8]);
    // un-permute: position 3 k1 + k2 holds X[k1 + 3 k2]
    cpx t;
    t = a[1]; a[1] = a[3]; a[3] = t;
    t = a[2]; a[2] = a[6]; a[6] = t;
    t = a[5]; a[5] = a[7]; a[7] = t;
}

template <int R, bool INV> struct DftR;
template <bool INV> struct DftR<8, INV> { ADM_HDM static void run(cpx* a) { dft8<INV>(a); } };
template <bool INV> struct DftR<9, INV> { ADM_HDM static void run(cpx* a) { dft9<INV>(a); } };

// ---- transcendental helpers ---------------------------------------------------------------
// sin/cos: short Taylor series for |x| < 0.5 (abs error < 2e-11, well below fp32 rounding), library
// sincosf otherwise.  The multislice phase k1*delta per slice is ~1e-4..1e-1 rad in every named config.
ADM_DEV void adm_sincos(float x, float* s, float* c) {
    if (fabsf(x) < 0.5f) {
        float x2 = x * x;
        float sp = fmaf(x2, 2.7557319e-6f, -1.9841270e-4f);
        sp = fmaf(sp, x2, 8.3333333e-3f);
        sp = fmaf(sp, x2, -1.6666667e-1f);
        *s = fmaf(sp * x2, x, x);
        float cp = fmaf(x2, -2.7557319e-7f, 2.4801587e-5f);
        cp = fmaf(cp, x2, -1.3888889e-3f);
        cp = fmaf(cp, x2, 4.1666667e-2f);
        cp = fmaf(cp, x2, -0.5f);
        *c = fmaf(cp, x2, 1.0f);
    } else {
#ifndef ADM_HOST_EMU
        sincosf(x, s, c);
#else
        *s = sinf(x); *c = cosf(x);
#endif
    }
}
