// Shared-memory mixed-radix FFT passes (P = 8 * R2, R2 = 8 or 9) used by the multislice and probe-shift kernels.
// See multislice.cu for the description of the pass structure and the permuted spectrum order.
#pragma once
#include "adm_common.cuh"

template <int P_, int R2_, int Q_> struct MsCfg {
    static constexpr int P = P_;
    static constexpr int R1 = 8;
    static constexpr int R2 = R2_;
    static constexpr int Q = Q_;            // thread rows: thread t -> (q = t / P, l = t % P)
    static constexpr int S = P_ + 1;        // psi row stride (complex)
    static constexpr int NACT = P_ * Q_;    // active threads
    static constexpr int NTHREADS = ((P_ * Q_ + 31) / 32) * 32;
    static constexpr size_t SMEM = sizeof(cpx) * (size_t)(P_ * (P_ + 1) + P_ * P_ + P_);
    static_assert(P_ == 8 * R2_, "P = 8 * R2");
};

// ---- generic passes over the shared-memory wavefield -----------------------------------------------
// y-direction passes: lanes along x (l = column).  x-direction passes: lanes along y (l = row).
template <class C, bool XDIR> ADM_DEV int addr(int l, int along) {
    return XDIR ? l * C::S + along : along * C::S + l;
}

// pass A (forward radix-8 + twiddle) along one direction
template <class C, bool XDIR> ADM_DEV void pass_A_fwd(cpx* psi, const cpx* tw, int q, int l) {
#pragma unroll
    for (int gi = 0; gi < (C::R2 + C::Q - 1) / C::Q; ++gi) {
        const int g = q + gi * C::Q;
        if ((C::R2 % C::Q) != 0 && g >= C::R2) break;
        cpx v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = psi[addr<C, XDIR>(l, C::R2 * j + g)];
        dft8<false>(v);
#pragma unroll
        for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw[g * k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) psi[addr<C, XDIR>(l, C::R2 * k + g)] = v[k];
    }
}
// pass A^-1 (inverse radix-8), no twiddle (applied by B^-1)
template <class C, bool XDIR> ADM_DEV void pass_A_inv(cpx* psi, int q, int l) {
#pragma unroll
    for (int gi = 0; gi < (C::R2 + C::Q - 1) / C::Q; ++gi) {
        const int g = q + gi * C::Q;
        if ((C::R2 % C::Q) != 0 && g >= C::R2) break;
        cpx v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = psi[addr<C, XDIR>(l, C::R2 * j + g)];
        dft8<true>(v);
#pragma unroll
        for (int k = 0; k < 8; ++k) psi[addr<C, XDIR>(l, C::R2 * k + g)] = v[k];
    }
}
// pass B forward (radix-R2 over the contiguous index)
template <class C, bool XDIR> ADM_DEV void pass_B_fwd(cpx* psi, int q, int l) {
#pragma unroll
    for (int gi = 0; gi < (C::R1 + C::Q - 1) / C::Q; ++gi) {
        const int g = q + gi * C::Q;
        if ((C::R1 % C::Q) != 0 && g >= C::R1) break;
        cpx v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[addr<C, XDIR>(l, C::R2 * g + n)];
        DftR<C::R2, false>::run(v);
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[addr<C, XDIR>(l, C::R2 * g + n)] = v[n];
    }
}
// pass B^-1 (inverse radix-R2 + conjugate twiddle)
template <class C, bool XDIR> ADM_DEV void pass_B_inv(cpx* psi, const cpx* tw, int q, int l) {
#pragma unroll
    for (int gi = 0; gi < (C::R1 + C::Q - 1) / C::Q; ++gi) {
        const int g = q + gi * C::Q;
        if ((C::R1 % C::Q) != 0 && g >= C::R1) break;
        cpx v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[addr<C, XDIR>(l, C::R2 * g + n)];
        DftR<C::R2, true>::run(v);
#pragma unroll
        for (int n = 1; n < C::R2; ++n) v[n] = cmulc(v[n], tw[g * n]);
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[addr<C, XDIR>(l, C::R2 * g + n)] = v[n];
    }
}
// P4: pass B forward, multiply by the prepared transfer function, pass B^-1 (x direction, lanes along y)
template <class C> ADM_DEV void pass_BHB_x(cpx* psi, const cpx* tw, const cpx* __restrict__ H, int q, int l) {
#pragma unroll
    for (int gi = 0; gi < (C::R1 + C::Q - 1) / C::Q; ++gi) {
        const int g = q + gi * C::Q;
        if ((C::R1 % C::Q) != 0 && g >= C::R1) break;
        cpx v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[l * C::S + C::R2 * g + n];
        DftR<C::R2, false>::run(v);
#pragma unroll
        for (int k = 0; k < C::R2; ++k) v[k] = cmul(v[k], H[(C::R2 * g + k) * C::P + l]);
        DftR<C::R2, true>::run(v);
#pragma unroll
        for (int n = 1; n < C::R2; ++n) v[n] = cmulc(v[n], tw[g * n]);
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[l * C::S + C::R2 * g + n] = v[n];
    }
}

// ---- separable transfer function H(ky, kx) = hy(ky) hx(kx) (Fresnel kernel, propagate.py:72) -----------------------
// hys / hxs are P-entry shared-memory tables in the permuted spectrum order (hxs carries the 1/P^2 scale); every lane
// of a warp reads the same entry (broadcast), so H costs no global / L1 traffic in the slice loop.
// pass B forward along y with the hy factor applied to the outputs
template <class C> ADM_DEV void pass_B_fwd_y_scaled(cpx* psi, const cpx* hys, int q, int l) {
#pragma unroll
    for (int gi = 0; gi < (C::R1 + C::Q - 1) / C::Q; ++gi) {
        const int g = q + gi * C::Q;
        if ((C::R1 % C::Q) != 0 && g >= C::R1) break;
        cpx v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[(C::R2 * g + n) * C::S + l];
        DftR<C::R2, false>::run(v);
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[(C::R2 * g + n) * C::S + l] = cmul(v[n], hys[C::R2 * g + n]);
    }
}
// P4 with the hx factor
template <class C> ADM_DEV void pass_BHB_x_sep(cpx* psi, const cpx* tw, const cpx* hxs, int q, int l) {
#pragma unroll
    for (int gi = 0; gi < (C::R1 + C::Q - 1) / C::Q; ++gi) {
        const int g = q + gi * C::Q;
        if ((C::R1 % C::Q) != 0 && g >= C::R1) break;
        cpx v[C::R2];
#pragma unroll
        for (int n = 0; n < C::R2; ++n) v[n] = psi[l * C::S + C::R2 * g + n];
        DftR<C::R2, false>::run(v);
#pragma unroll
        for (int k = 0; k < C::R2; ++k) v[k] = cmul(v[k], hxs[C::R2 * g + k]);
        DftR<C::R2, true>::run(v);
#pragma unroll
        for (int n = 1; n < C::R2; ++n) v[n] = cmulc(v[n], tw[g * n]);
#pragma unroll
        for (int n = 0; n < C::R2; ++n) psi[l * C::S + C::R2 * g + n] = v[n];
    }
}

template <class C> ADM_DEV int perm_freq(int p) { return p / C::R2 + C::R1 * (p % C::R2); }

template <class C> ADM_DEV void init_twiddles(cpx* tw, int t) {
    if (t < C::P) {
#ifndef ADM_HOST_EMU
        float s, c;
        sincospif(2.0f * (float)t / (float)C::P, &s, &c);
        tw[t] = mk(c, -s);
#else
        double a = 2.0 * 3.14159265358979323846 * t / C::P;
        tw[t] = mk((float)cos(a), (float)-sin(a));
#endif
    }
}

// P2..P6 : psi <- IFFT2(H . FFT2(psi)) given that pass A along y has already been applied, and leaving the
// final pass A^-1 along y to the caller (fused with the next modulation).
template <class C> ADM_DEV void propagate_core(cpx* psi, const cpx* tw, const cpx* H, int q, int l, bool active) {
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


// same with a separable transfer function: hs = [hy(P) | hx(P)] in shared memory
template <class C> ADM_DEV void propagate_core_sep(cpx* psi, const cpx* tw, const cpx* hs, int q, int l, bool active) {
    if (active) pass_B_fwd_y_scaled<C>(psi, hs, q, l);
    __syncthreads();
    if (active) pass_A_fwd<C, true>(psi, tw, q, l);
    __syncthreads();
    if (active) pass_BHB_x_sep<C>(psi, tw, hs + C::P, q, l);
    __syncthreads();
    if (active) pass_A_inv<C, true>(psi, q, l);
    __syncthreads();
    if (active) pass_B_inv<C, false>(psi, tw, q, l);
    __syncthreads();
}

typedef MsCfg<72, 9, 3> Cfg72;
typedef MsCfg<64, 8, 4> Cfg64;
// "wide" shapes: one butterfly group per thread (three times the warps, a third of the work per thread)
typedef MsCfg<72, 9, 9> Cfg72w;
typedef MsCfg<64, 8, 8> Cfg64w;

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
