// Packed complex arithmetic: one complex value = one 64-bit register pair (re, im), processed by the sm_100a packed
// fp32x2 instructions (FADD2 / FMUL2 / FFMA2).  The SASS operands of these instructions take a half-swap (.LO_HI), a
// per-half sign pattern (.NP / .PN) and a scalar broadcast (.F32) for free, and ptxas folds an explicit
// unpack -> negate -> pack sequence into those modifiers (checked with cuobjdump: a complex multiply is FMUL2 + FFMA2,
// a +-i rotation is an operand modifier of the consuming FADD2).  The interleaved (re, im) layout of shared and global
// memory is exactly the register-pair layout, so an LDS.64 feeds the packed butterflies directly: no re-pairing, no
// layout change, and the complex butterflies of the FFT cost 0.41x (radix 8) / 0.44x (radix 9) of the scalar
// instruction count.  The FP32 lane rate is unchanged (tools/f32x2_probe.cu); what is saved is issue slots.
//
// Host emulation (-DADM_HOST_EMU): cpk is two floats and the primitives are plain C, so the same butterfly code is
// unit-tested on the CPU (tests/test_kernel_logic_emu.py).
#pragma once
#include "adm_common.cuh"

#ifndef ADM_HOST_EMU
struct cpk {
    unsigned long long v;
};
ADM_DEV cpk kmk(float x, float y) { cpk r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y)); return r; }
ADM_DEV void kun(cpk a, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); }
ADM_DEV cpk operator+(cpk a, cpk b) { cpk r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
ADM_DEV cpk operator-(cpk a, cpk b) { cpk r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
// a * s and a * s + c with the scalar s broadcast to both halves
ADM_DEV cpk kscale(cpk a, float s) {
    cpk r, b = kmk(s, s);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
ADM_DEV cpk kfmas(cpk a, float s, cpk c) {
    cpk r, b = kmk(s, s);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
#else
struct cpk {
    float x, y;
};
static inline cpk kmk(float x, float y) { cpk r; r.x = x; r.y = y; return r; }
static inline void kun(cpk a, float& x, float& y) { x = a.x; y = a.y; }
static inline cpk operator+(cpk a, cpk b) { return kmk(a.x + b.x, a.y + b.y); }
static inline cpk operator-(cpk a, cpk b) { return kmk(a.x - b.x, a.y - b.y); }
static inline cpk kscale(cpk a, float s) { return kmk(a.x * s, a.y * s); }
static inline cpk kfmas(cpk a, float s, cpk c) { return kmk(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y)); }
#endif

ADM_DEV float cx(cpk a) { float x, y; kun(a, x, y); return x; }
ADM_DEV float cy(cpk a) { float x, y; kun(a, x, y); return y; }
ADM_DEV float cx(cpx a) { return a.x; }
ADM_DEV float cy(cpx a) { return a.y; }
template <class T> struct is_cpk { static constexpr bool v = false; };
template <> struct is_cpk<cpk> { static constexpr bool v = true; };
template <class T> ADM_DEV T tmk(float x, float y) {
    if constexpr (is_cpk<T>::v) return kmk(x, y);
    else return mk(x, y);
}

// -i a = (a.y, -a.x) and +i a = (-a.y, a.x): folded by ptxas into the operand modifiers of the consumer
ADM_DEV cpk kswp_pn(cpk a) { float x, y; kun(a, x, y); return kmk(y, -x); }
ADM_DEV cpk kswp_np(cpk a) { float x, y; kun(a, x, y); return kmk(-y, x); }
template <bool INV> ADM_DEV cpk kmi(cpk a) { return INV ? kswp_np(a) : kswp_pn(a); }       // (-+i) a
template <bool INV> ADM_DEV cpk add_mi(cpk a, cpk w) { return a + kmi<INV>(w); }
template <bool INV> ADM_DEV cpk sub_mi(cpk a, cpk w) { return a + kmi<!INV>(w); }

#ifndef ADM_HOST_EMU
ADM_DEV cpk cconj(cpk a) {      // explicit FMUL2 by (1, -1)
    cpk r, b = kmk(1.0f, -1.0f);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
#else
static inline cpk cconj(cpk a) { return kmk(a.x, -a.y); }
#endif
ADM_DEV cpk cscale(cpk a, float s) { return kscale(a, s); }
ADM_DEV cpk cmul(cpk a, cpk b) { float bx, by; kun(b, bx, by); return kfmas(kswp_np(a), by, kscale(a, bx)); }
// a * conj(b)
ADM_DEV cpk cmulc(cpk a, cpk b) { float bx, by; kun(b, bx, by); return kfmas(kswp_pn(a), by, kscale(a, bx)); }
// psi (1 + T') and g conj(1 + T')
ADM_DEV cpk modulate(cpk v, cpk tp) { float tx, ty; kun(tp, tx, ty); return kfmas(kswp_np(v), ty, kfmas(v, tx, v)); }
ADM_DEV cpk modulate_conj(cpk v, cpk tp) { float tx, ty; kun(tp, tx, ty); return kfmas(kswp_pn(v), ty, kfmas(v, tx, v)); }

// ---- butterflies (same conventions as dft8 / dft9 of adm_common.cuh) ------------------------------------------------
template <bool INV> ADM_DEV void dft8(cpk* a) {
    cpk b0 = a[0] + a[4], c0 = a[0] - a[4];
    cpk b1 = a[1] + a[5], c1 = a[1] - a[5];
    cpk b2 = a[2] + a[6], c2 = a[2] - a[6];
    cpk b3 = a[3] + a[7], c3 = a[3] - a[7];
    {   // even outputs: 4-point DFT of b
        cpk e0 = b0 + b2, e1 = b0 - b2, e2 = b1 + b3, t = b1 - b3;
        a[0] = e0 + e2; a[4] = e0 - e2; a[2] = add_mi<INV>(e1, t); a[6] = sub_mi<INV>(e1, t);
    }
    {   // odd outputs: 4-point DFT of (c0, W8 c1, -+i c2, W8^3 c3) with W8 c1 = s r1, W8^3 c3 = -s r3 (s = 1/sqrt 2)
        cpk r1 = add_mi<INV>(c1, c1), r3 = add_mi<!INV>(c3, c3);
        cpk e0 = add_mi<INV>(c0, c2), e1 = sub_mi<INV>(c0, c2);
        cpk e2 = r1 - r3, t = r1 + r3;
        a[1] = kfmas(e2, ADM_SQRT1_2, e0);
        a[5] = kfmas(e2, -ADM_SQRT1_2, e0);
        a[3] = kfmas(kmi<INV>(t), ADM_SQRT1_2, e1);
        a[7] = kfmas(kmi<!INV>(t), ADM_SQRT1_2, e1);
    }
}

template <bool INV> ADM_DEV void dft4(cpk& d0, cpk& d1, cpk& d2, cpk& d3) {
    cpk e0 = d0 + d2, e1 = d0 - d2, e2 = d1 + d3, t = d1 - d3;
    d0 = e0 + e2; d2 = e0 - e2; d1 = add_mi<INV>(e1, t); d3 = sub_mi<INV>(e1, t);
}

template <bool INV> ADM_DEV void dft3(cpk& a0, cpk& a1, cpk& a2) {
    cpk t = a1 + a2;
    cpk u = kfmas(t, -0.5f, a0);
    cpk d = a1 - a2;
    a0 = a0 + t;
    a1 = kfmas(kmi<INV>(d), ADM_SQRT3_2, u);       // u -+ i c d
    a2 = kfmas(kmi<!INV>(d), ADM_SQRT3_2, u);
}

// multiply by exp(-+ i phi), (c, s) = (cos phi, sin phi)
template <bool INV> ADM_DEV cpk tw9(cpk a, float c, float s) { return kfmas(kmi<INV>(a), s, kscale(a, c)); }

template <bool INV> ADM_DEV void dft9(cpk* a) {
    dft3<INV>(a[0], a[3], a[6]);
    dft3<INV>(a[1], a[4], a[7]);
    dft3<INV>(a[2], a[5], a[8]);
    a[4] = tw9<INV>(a[4], ADM_C9_1, ADM_S9_1);
    a[5] = tw9<INV>(a[5], ADM_C9_2, ADM_S9_2);
    a[7] = tw9<INV>(a[7], ADM_C9_2, ADM_S9_2);
    a[8] = tw9<INV>(a[8], ADM_C9_4, ADM_S9_4);
    dft3<INV>(a[0], a[1], a[2]);
    dft3<INV>(a[3], a[4], a[5]);
    dft3<INV>(a[6], a[7], a[8]);
    cpk t;
    t = a[1]; a[1] = a[3]; a[3] = t;
    t = a[2]; a[2] = a[6]; a[6] = t;
    t = a[5]; a[5] = a[7]; a[7] = t;
}

template <int R, bool INV, class T> ADM_DEV void dftr(T* a) {
    if constexpr (R == 8) dft8<INV>(a);
    else dft9<INV>(a);
}
