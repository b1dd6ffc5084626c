// Shared-memory mixed-radix FFT passes (P = 8 * R2, R2 = 8 or 9) used by the multislice and probe-shift kernels.
// See multislice.cu for the description of the pass structure and the permuted spectrum order.
#pragma once
#include "adm_common.cuh"
#include "cpk.cuh"

template <int P_, int R2_, int Q_, bool PFA_ = false> struct MsCfg {
    static constexpr int P = P_;
    static constexpr int R1 = 8;
    static constexpr int R2 = R2_;
    static constexpr int Q = Q_;            // thread rows: thread t -> (q = t / P, l = t % P)
    static constexpr int S = P_ + 1;        // psi row stride (complex)
    static constexpr int NACT = P_ * Q_;    // active threads
    static constexpr int NTHREADS = ((P_ * Q_ + 31) / 32) * 32;
    static constexpr int NGA = (R2_ + Q_ - 1) / Q_;   // A-type groups per thread (radix-8 butterflies, R2 groups per line)
    static constexpr int NGB = (8 + Q_ - 1) / Q_;     // B-type groups per thread (radix-R2 butterflies, 8 groups per line)
    // Good-Thomas prime-factor indexing (8 and 9 are coprime): NO twiddle factors between the radix-8 and radix-9
    // stages; only the index maps change (see Idx below).  P = 64 = 8 * 8 keeps Cooley-Tukey + twiddles.
    static constexpr bool PFA = PFA_;
    static_assert(P_ == 8 * R2_, "P = 8 * R2");
    static_assert(!PFA_ || (R2_ == 9 && Q_ == 3), "prime-factor path: P = 72, three thread rows");
};

// ---- index maps -------------------------------------------------------------------------------------------------
// Along one axis a line of P points is processed by "A" butterflies (radix 8 over a, for fixed b) and "B" butterflies
// (radix R2 over b, for fixed a); thread row q owns the groups q, q + Q, q + 2Q, ...
//   Cooley-Tukey : position(a, b) = R2 a + b                (real space natural, spectrum X[a + 8 b] at R2 a + b)
//   prime factor : position(a, b) = (9 a + 8 b) mod 72      (real space natural, spectrum X[(9 a + 64 b) mod 72])
// For the prime-factor map the position of element e of group gi is c(e, gi) + r(q) (mod 72) with a compile-time c
// and r in {0, u, 2u} (u = 8 for A, 9 for B): the wrap-around is folded into three per-thread offsets selected at
// compile time by the range of c, so the unrolled butterflies address shared memory with immediate offsets only.
template <class C> struct Idx {
    int q, oa0, oa1, oa2, ob0, ob1, ob2;
    ADM_HDM explicit Idx(int q_) : q(q_), oa0(0), oa1(0), oa2(0), ob0(0), ob1(0), ob2(0) {
        if constexpr (C::PFA) {
            const int ra = 8 * q, rb = 9 * q;
            oa0 = ra; oa1 = ra - (ra >= 16 ? 72 : 0); oa2 = ra - (ra >= 8 ? 72 : 0);
            ob0 = rb; ob1 = rb - (rb >= 18 ? 72 : 0); ob2 = rb - (rb >= 9 ? 72 : 0);
        }
    }
    // group owned by this thread (b for A-type, a for B-type)
    ADM_HDM int group(int gi) const { return q + C::Q * gi; }
    // position of element j (0..7) of the gi-th A-type group
    ADM_HDM int a(int j, int gi) const {
        if constexpr (C::PFA) {
            const int c = (9 * j + 8 * C::Q * gi) % 72;
            return c + (c < 56 ? oa0 : (c < 64 ? oa1 : oa2));
        } else {
            return C::R2 * j + (q + C::Q * gi);
        }
    }
    // position of element n (0..R2-1) of the gi-th B-type group
    ADM_HDM int b(int n, int gi) const {
        if constexpr (C::PFA) {
            const int c = (8 * n + 9 * C::Q * gi) % 72;
            return c + (c < 54 ? ob0 : (c < 63 ? ob1 : ob2));
        } else {
            return C::R2 * (q + C::Q * gi) + n;
        }
    }
};

// frequency index held at spectrum position p
template <class C> ADM_DEV int perm_freq(int p) {
    if constexpr (C::PFA) {
        const int k1 = p % 8, k2 = (9 - p % 9) % 9;
        return (9 * k1 + 64 * k2) % 72;
    } else {
        return p / C::R2 + C::R1 * (p % C::R2);
    }
}

// ---- generic passes over the shared-memory wavefield -----------------------------------------------
// y-direction passes: lanes along x (l = column).  x-direction passes: lanes along y (l = row).
template <class C, bool XDIR> ADM_DEV int addr(int l, int along) {
    return XDIR ? l * C::S + along : along * C::S + l;
}
#define ADM_FOR_GROUPS_A(gi) _Pragma("unroll") for (int gi = 0; gi < C::NGA; ++gi) if ((C::R2 % C::Q) == 0 || I.group(gi) < C::R2)
#define ADM_FOR_GROUPS_B(gi) _Pragma("unroll") for (int gi = 0; gi < C::NGB; ++gi) if ((C::R1 % C::Q) == 0 || I.group(gi) < C::R1)

// pass A (forward radix-8 [+ twiddle]) along one direction
template <class C, bool XDIR, class T> ADM_DEV void pass_A_fwd(T* psi, const T* tw, int q, int l) {
    const Idx<C> I(q);
    ADM_FOR_GROUPS_A(gi) {
        T v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = psi[addr<C, XDIR>(l, I.a(j, gi))];
        dft8<false>(v);
        if constexpr (!C::PFA) {
#pragma unroll
            for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw[I.group(gi) * k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) psi[addr<C, XDIR>(l, I.a(k, gi))] = v[k];
    }
}
// pass A^-1 (inverse radix-8), no twiddle (applied by B^-1)
template <class C, bool XDIR, class T> ADM_DEV void pass_A_inv(T* psi, int q, int l) {
    const Idx<C> I(q);
    ADM_FOR_GROUPS_A(gi) {
        T v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = psi[addr<C, XDIR>(l, I.a(j, gi))];
        dft8<true>(v);
#pragma unroll
        for (int k = 0; k < 8; ++k) psi[addr<C, XDIR>(l, I.a(k, gi))] = v[k];
    }
}
// pass B forward (radix-R2)
template <class C, bool XDIR, class T> ADM_DEV void pass_B_fwd(T* psi, int q, int l) {
    const Idx<C> I(q);
    ADM_FOR_GROUPS_B(gi) {
        T v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[addr<C, XDIR>(l, I.b(n, gi))];
        dftr<C::R2, false>(v);
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[addr<C, XDIR>(l, I.b(n, gi))] = v[n];
    }
}
// pass B^-1 (inverse radix-R2 [+ conjugate twiddle])
template <class C, bool XDIR, class T> ADM_DEV void pass_B_inv(T* psi, const T* tw, int q, int l) {
    const Idx<C> I(q);
    ADM_FOR_GROUPS_B(gi) {
        T v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[addr<C, XDIR>(l, I.b(n, gi))];
        dftr<C::R2, true>(v);
        if constexpr (!C::PFA) {
#pragma unroll
            for (int n = 1; n < C::R2; ++n) v[n] = cmulc(v[n], tw[I.group(gi) * n]);
        }
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[addr<C, XDIR>(l, I.b(n, gi))] = v[n];
    }
}
// P4: pass B forward, multiply by the prepared transfer function, pass B^-1 (x direction, lanes along y)
template <class C, class T> ADM_DEV void pass_BHB_x(T* psi, const T* tw, const T* __restrict__ H, int q, int l) {
    const Idx<C> I(q);
    ADM_FOR_GROUPS_B(gi) {
        T v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[l * C::S + I.b(n, gi)];
        dftr<C::R2, false>(v);
#pragma unroll
        for (int k = 0; k < C::R2; ++k) v[k] = cmul(v[k], H[I.b(k, gi) * C::P + l]);
        dftr<C::R2, true>(v);
        if constexpr (!C::PFA) {
#pragma unroll
            for (int n = 1; n < C::R2; ++n) v[n] = cmulc(v[n], tw[I.group(gi) * n]);
        }
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[l * C::S + I.b(n, gi)] = v[n];
    }
}

// ---- shared-memory layout of the separable transfer-function tables ------------------------------------------------
// A thread multiplies the R2 outputs of its gi-th B-type group by h[I.b(k, gi)], k = 0 .. R2-1: positions that depend on
// the thread row q only.  The tables are therefore stored per (thread row, group): hsq[(q NGB + gi) RP + k] = h[I.b(k, gi)],
// RP = R2 rounded up to even, so that one 16-byte broadcast load (LDS.128) brings two factors: 5 instead of 9 shared-memory
// wavefronts per group for R2 = 9 (the table reads were 8 % of the wavefronts of a slice).
template <class C> struct HsLayout {
    static constexpr int RP = (C::R2 + 1) & ~1;
    static constexpr int N = C::Q * C::NGB * RP;      // entries per direction
};
// position along the axis whose factor sits at table entry i (-1: padding)
template <class C> ADM_DEV int hs_source(int i) {
    constexpr int RP = HsLayout<C>::RP;
    const int q = i / (C::NGB * RP), gi = (i / RP) % C::NGB, k = i % RP;
    if (k >= C::R2 || q + C::Q * gi >= C::R1) return -1;
    return Idx<C>(q).b(k, gi);
}
// fill both tables (hy | hx) from the prepared [2][P] vectors in global memory; threads t < 2 N
template <class C, class T> ADM_DEV void hs_fill(T* hs, const T* Hsep, int t) {
    constexpr int N = HsLayout<C>::N;
    if (t < 2 * N) {
        const int src = hs_source<C>(t % N);
        hs[t] = src >= 0 ? Hsep[(t / N) * C::P + src] : tmk<T>(0.f, 0.f);
    }
}
template <class T> ADM_DEV void lds_pair(const T* p, T& a, T& b) {
#ifndef ADM_HOST_EMU
    const float4 f = *reinterpret_cast<const float4*>(p);
    a = tmk<T>(f.x, f.y); b = tmk<T>(f.z, f.w);
#else
    a = p[0]; b = p[1];
#endif
}
// v[k] *= h[I.b(k, gi)] for the R2 outputs of the gi-th B-type group of thread row q
template <class C, class T> ADM_DEV void hs_apply_group(const T* hsq, int q, int gi, T* v) {
    constexpr int RP = HsLayout<C>::RP;
    const T* hq = hsq + (q * C::NGB + gi) * RP;
#pragma unroll
    for (int k = 0; k < RP; k += 2) {
        T ha, hb;
        lds_pair(hq + k, ha, hb);
        v[k] = cmul(v[k], ha);
        if (k + 1 < C::R2) v[k + 1] = cmul(v[k + 1], hb);
    }
}

// ---- separable transfer function, FIVE passes per slice ------------------------------------------------------------
// With H(ky, kx) = hy(ky) hx(kx) (the paraxial Fresnel kernel, propagate.py:72; hys / hxs are P-entry shared-memory
// tables indexed by spectrum position, hxs carries the 1/P^2 scale) the propagation is a 1-D convolution along y followed by one along x, and the two
// commute with each other's transforms.  After P1 (which ends with the radix-8 stage Ay) the y convolution therefore
// completes in ONE pass [By . hy . By^-1] (the closing Ay^-1 is the head of the next slice's P1), and the x convolution
// is [Ax] [Bx . hx . Bx^-1] [Ax^-1]: five shared-memory round trips and five barriers per slice instead of six.
// P2': y direction, lanes along x; hys indexed by the spectrum position along y (9 broadcast table reads per group).
template <class C, class T> ADM_DEV void pass_BHB_y_sep(T* psi, const T* tw, const T* hys, int q, int l) {
    const Idx<C> I(q);
    ADM_FOR_GROUPS_B(gi) {
        T v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[I.b(n, gi) * C::S + l];
        dftr<C::R2, false>(v);
        hs_apply_group<C>(hys, q, gi, v);
        dftr<C::R2, true>(v);
        if constexpr (!C::PFA) {
#pragma unroll
            for (int n = 1; n < C::R2; ++n) v[n] = cmulc(v[n], tw[I.group(gi) * n]);
        }
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[I.b(n, gi) * C::S + l] = v[n];
    }
}
// P4': x direction, lanes along y; only the x factor (hxs carries the 1/P^2 scale)
template <class C, class T> ADM_DEV void pass_BHB_x_sep1(T* psi, const T* tw, const T* hxs, int q, int l) {
    const Idx<C> I(q);
    ADM_FOR_GROUPS_B(gi) {
        T v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[l * C::S + I.b(n, gi)];
        dftr<C::R2, false>(v);
        hs_apply_group<C>(hxs, q, gi, v);
        dftr<C::R2, true>(v);
        if constexpr (!C::PFA) {
#pragma unroll
            for (int n = 1; n < C::R2; ++n) v[n] = cmulc(v[n], tw[I.group(gi) * n]);
        }
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[l * C::S + I.b(n, gi)] = v[n];
    }
}
// P2' .. P5 : psi <- IFFT2(H . FFT2(psi)) for separable H, given that pass A along y has already been applied and leaving
// the final pass A^-1 along y to the caller.  hs = [hy(P) | hx(P)] in shared memory.
template <class C, class T> ADM_DEV void propagate_core_sep5(T* psi, const T* tw, const T* hs, int q, int l, bool active) {
    if (active) pass_BHB_y_sep<C>(psi, tw, hs, q, l);
    __syncthreads();
    if (active) pass_A_fwd<C, true>(psi, tw, q, l);
    __syncthreads();
    if (active) pass_BHB_x_sep1<C>(psi, tw, hs + HsLayout<C>::N, q, l);
    __syncthreads();
    if (active) pass_A_inv<C, true>(psi, q, l);
    __syncthreads();
}

template <class C, class T> ADM_DEV void init_twiddles(T* tw, int t) {
    if (t < C::P) {
#ifndef ADM_HOST_EMU
        float s, c;
        sincospif(2.0f * (float)t / (float)C::P, &s, &c);
        tw[t] = tmk<T>(c, -s);
#else
        double a = 2.0 * 3.14159265358979323846 * t / C::P;
        tw[t] = tmk<T>((float)cos(a), (float)-sin(a));
#endif
    }
}

// P2..P6 : psi <- IFFT2(H . FFT2(psi)) given that pass A along y has already been applied, and leaving the
// final pass A^-1 along y to the caller (fused with the next modulation).
template <class C, class T> ADM_DEV void propagate_core(T* psi, const T* tw, const T* H, int q, int l, bool active) {
    if (active) pass_B_fwd<C, false>(psi, q, l);
    __syncthreads();
    if (active) pass_A_fwd<C, true>(psi, tw, q, l);
    __syncthreads();
    if (active) pass_BHB_x<C>(psi, tw, H, q, l);
    __syncthreads();
    if (active) pass_A_inv<C, true>(psi, q, l);
    __syncthreads();
    if (active) pass_B_inv<C, false>(psi, tw, q, l);
    __syncthreads();
}


typedef MsCfg<72, 9, 3, true> Cfg72;     // prime-factor indexing, no twiddles
typedef MsCfg<72, 9, 3, false> Cfg72ct;  // Cooley-Tukey variant of the same shape (tuning / cross-check)
typedef MsCfg<64, 8, 4> Cfg64;

#ifndef ADM_HOST_EMU
#define ADM_LAUNCH(kern, grid, block, smem, stream, ...)                                         \
    do {                                                                                         \
        kern<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__);                        \
        if (cudaGetLastError() != cudaSuccess) return ADM_ERR_CUDA;                              \
    } while (0)
template <class K> static int set_smem(K kern, size_t bytes) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : 1;
}
#define ADM_SMEM_DECL extern __shared__ __align__(128) unsigned char smem_raw[];
#else
#define ADM_LAUNCH(kern, grid, block, smem, stream, ...) emu_launch(dim3(grid), dim3(block), smem, [=] { kern(__VA_ARGS__); })
template <class K> static int set_smem(K, size_t) { return 0; }
#define ADM_SMEM_DECL unsigned char* smem_raw = (unsigned char*)emu_smem;
#endif
