// Packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2) and radix-8 / radix-9 butterflies on PAIRS of complex
// values: cx2 = (re_A, re_B | im_A, im_B) holds the same element of two independent wavefields A and B, so every
// butterfly instruction works on both at once.  Measured on B200 (tools/f32x2_probe.cu): the packed instructions do not
// raise the FP32 rate (128 lanes/clk/SM either way) but halve the ISSUE slots, which is what the shared-memory FFT
// kernels are short of.  ptxas folds broadcast constants into the instruction (FMUL2 R, R.F32x2, imm) and negations
// into operand modifiers, so the butterflies below are written with add / sub / multiply-by-scalar only.
//
// Host emulation (-DADM_HOST_EMU): f2 is two floats and the operations are plain C.
#pragma once
#include "adm_common.cuh"

#ifndef ADM_HOST_EMU
struct f2 {
    unsigned long long v;
};
ADM_DEV f2 pack2(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
ADM_DEV void unpack2(f2 a, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); }
ADM_DEV f2 operator+(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
ADM_DEV f2 operator-(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
ADM_DEV f2 operator*(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
ADM_DEV f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
#else
struct f2 {
    float a, b;
};
static inline f2 pack2(float a, float b) { f2 r; r.a = a; r.b = b; return r; }
static inline void unpack2(f2 v, float& x, float& y) { x = v.a; y = v.b; }
static inline f2 operator+(f2 a, f2 b) { return pack2(a.a + b.a, a.b + b.b); }
static inline f2 operator-(f2 a, f2 b) { return pack2(a.a - b.a, a.b - b.b); }
static inline f2 operator*(f2 a, f2 b) { return pack2(a.a * b.a, a.b * b.b); }
static inline f2 fma2(f2 a, f2 b, f2 c) { return pack2(fmaf(a.a, b.a, c.a), fmaf(a.b, b.b, c.b)); }
#endif

// a * c and a * c + b with a scalar c broadcast to both halves
ADM_DEV f2 mulc(f2 a, float c) { return a * pack2(c, c); }
ADM_DEV f2 fmac(f2 a, float c, f2 b) { return fma2(a, pack2(c, c), b); }

struct alignas(16) cx2 {
    f2 x, y;      // (re_A, re_B), (im_A, im_B)
};
ADM_DEV cx2 mk2(f2 x, f2 y) { cx2 r; r.x = x; r.y = y; return r; }
ADM_DEV cx2 operator+(cx2 a, cx2 b) { return mk2(a.x + b.x, a.y + b.y); }
ADM_DEV cx2 operator-(cx2 a, cx2 b) { return mk2(a.x - b.x, a.y - b.y); }
ADM_DEV cx2 pack_cx(cpx a, cpx b) { return mk2(pack2(a.x, b.x), pack2(a.y, b.y)); }
ADM_DEV void unpack_cx(cx2 v, cpx& a, cpx& b) { unpack2(v.x, a.x, b.x); unpack2(v.y, a.y, b.y); }
ADM_DEV cx2 conj2(cx2 a) { return mk2(a.x, mulc(a.y, -1.0f)); }
ADM_DEV cx2 scale2(cx2 a, float s) { return mk2(mulc(a.x, s), mulc(a.y, s)); }
// multiply both wavefields by the same complex scalar h
ADM_DEV cx2 cmul2s(cx2 v, cpx h) { return mk2(fmac(v.y, -h.y, mulc(v.x, h.x)), fmac(v.y, h.x, mulc(v.x, h.y))); }

// a + (-i) w (forward) / a + (+i) w (inverse), and the matching differences: no negations needed
template <bool INV> ADM_DEV cx2 add_mi(cx2 a, cx2 w) { return INV ? mk2(a.x - w.y, a.y + w.x) : mk2(a.x + w.y, a.y - w.x); }
template <bool INV> ADM_DEV cx2 sub_mi(cx2 a, cx2 w) { return INV ? mk2(a.x + w.y, a.y - w.x) : mk2(a.x - w.y, a.y + w.x); }

// 8-point DFT, natural in / natural out (INV: conjugate twiddles, unnormalised) -- same convention as dft8 in
// adm_common.cuh
template <bool INV> ADM_DEV void dft8_2(cx2* a) {
    cx2 b0 = a[0] + a[4], c0 = a[0] - a[4];
    cx2 b1 = a[1] + a[5], c1 = a[1] - a[5];
    cx2 b2 = a[2] + a[6], c2 = a[2] - a[6];
    cx2 b3 = a[3] + a[7], c3 = a[3] - a[7];
    // even outputs: 4-point DFT of b
    {
        cx2 e0 = b0 + b2, e1 = b0 - b2, e2 = b1 + b3, t = b1 - b3;
        a[0] = e0 + e2; a[4] = e0 - e2; a[2] = add_mi<INV>(e1, t); a[6] = sub_mi<INV>(e1, t);
    }
    // odd outputs: 4-point DFT of (c0, W8 c1, -+i c2, W8^3 c3); W8 c1 = s r1, W8^3 c3 = -s r3 with the unscaled
    // rotations r1, r3: the factor s = 1/sqrt 2 is folded into the last butterfly (FFMA2), as in dft8
    {
        cx2 r1, r3;
        if (!INV) { r1 = mk2(c1.x + c1.y, c1.y - c1.x); r3 = mk2(c3.x - c3.y, c3.x + c3.y); }
        else      { r1 = mk2(c1.x - c1.y, c1.x + c1.y); r3 = mk2(c3.x + c3.y, c3.y - c3.x); }
        cx2 e0 = add_mi<INV>(c0, c2), e1 = sub_mi<INV>(c0, c2);
        cx2 e2 = r1 - r3, t = r1 + r3;           // (W8 c1 + W8^3 c3) / s, (W8 c1 - W8^3 c3) / s
        a[1] = mk2(fmac(e2.x, ADM_SQRT1_2, e0.x), fmac(e2.y, ADM_SQRT1_2, e0.y));
        a[5] = mk2(fmac(e2.x, -ADM_SQRT1_2, e0.x), fmac(e2.y, -ADM_SQRT1_2, e0.y));
        if (!INV) {
            a[3] = mk2(fmac(t.y, ADM_SQRT1_2, e1.x), fmac(t.x, -ADM_SQRT1_2, e1.y));
            a[7] = mk2(fmac(t.y, -ADM_SQRT1_2, e1.x), fmac(t.x, ADM_SQRT1_2, e1.y));
        } else {
            a[3] = mk2(fmac(t.y, -ADM_SQRT1_2, e1.x), fmac(t.x, ADM_SQRT1_2, e1.y));
            a[7] = mk2(fmac(t.y, ADM_SQRT1_2, e1.x), fmac(t.x, -ADM_SQRT1_2, e1.y));
        }
    }
}

template <bool INV> ADM_DEV void dft3_2(cx2& a0, cx2& a1, cx2& a2) {
    cx2 t = a1 + a2;
    cx2 u = mk2(fmac(t.x, -0.5f, a0.x), fmac(t.y, -0.5f, a0.y));
    cx2 d = a1 - a2;
    a0 = a0 + t;
    cx2 p = mk2(fmac(d.y, ADM_SQRT3_2, u.x), fmac(d.x, -ADM_SQRT3_2, u.y));
    cx2 q = mk2(fmac(d.y, -ADM_SQRT3_2, u.x), fmac(d.x, ADM_SQRT3_2, u.y));
    if (!INV) { a1 = p; a2 = q; } else { a1 = q; a2 = p; }
}

// multiply by exp(-+ i phi), (c, s) = (cos phi, sin phi)
template <bool INV> ADM_DEV cx2 tw9_2(cx2 a, float c, float s) {
    return INV ? mk2(fmac(a.y, -s, mulc(a.x, c)), fmac(a.x, s, mulc(a.y, c)))
               : mk2(fmac(a.y, s, mulc(a.x, c)), fmac(a.x, -s, mulc(a.y, c)));
}

// 9-point DFT, natural in / natural out -- same convention as dft9 in adm_common.cuh
template <bool INV> ADM_DEV void dft9_2(cx2* a) {
    dft3_2<INV>(a[0], a[3], a[6]);
    dft3_2<INV>(a[1], a[4], a[7]);
    dft3_2<INV>(a[2], a[5], a[8]);
    a[4] = tw9_2<INV>(a[4], ADM_C9_1, ADM_S9_1);
    a[5] = tw9_2<INV>(a[5], ADM_C9_2, ADM_S9_2);
    a[7] = tw9_2<INV>(a[7], ADM_C9_2, ADM_S9_2);
    a[8] = tw9_2<INV>(a[8], ADM_C9_4, ADM_S9_4);
    dft3_2<INV>(a[0], a[1], a[2]);
    dft3_2<INV>(a[3], a[4], a[5]);
    dft3_2<INV>(a[6], a[7], a[8]);
    cx2 t;
    t = a[1]; a[1] = a[3]; a[3] = t;
    t = a[2]; a[2] = a[6]; a[6] = t;
    t = a[5]; a[5] = a[7]; a[7] = t;
}
