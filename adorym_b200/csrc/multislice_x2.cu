// Multislice forward for P = 72, TWO wavefields per CTA on packed fp32x2 arithmetic (sm_100a).
//
// Same six-pass prime-factor FFT chain as multislice.cu (P1 [Ay^-1 . (1 + T') . Ay]  P2 By  P3 Ax  P4 [Bx . H . Bx^-1]
// P5 Ax^-1  P6 By^-1, see there), but a CTA carries wavefields 2 c and 2 c + 1 and every shared-memory element is the
// pair (re_A, re_B | im_A, im_B) (cx2, 16 bytes): each LDS.128 / STS.128 / FADD2 / FMUL2 / FFMA2 serves both
// wavefields, so the warp-instruction count per wavefield is halved (the six-pass kernels were bound by issue slots and
// shared-memory latency, not by the FP32 lanes: profiles/r01d).  The transmission tile is read with plain coalesced
// loads (lanes along x; the volume is L2 resident) at the top of P1, so the only shared memory is the paired wavefield
// (84 KB) and two CTAs stay resident per SM.  Checkpoints keep the layout of multislice.cu ([wf][Nz][P][P+2], column
// offset px & 1), so the adjoint kernels of either file consume them.
//
// Restrictions (otherwise multislice.cu serves the call): P = 72, separable in-object transfer function, detector
// transform 'far field' or none.  Reference behaviour reproduced: as multislice.cu.
#include <stdlib.h>
#include "fft_passes.cuh"
#include "f32x2.cuh"
#include "ms_params.cuh"

namespace msx2 {
typedef Cfg72 C;
constexpr int P = 72, R2 = 9, S = 73, TP = P + 2;
constexpr int NTHREADS = C::NTHREADS;                 // 224 (216 active): thread (q, l), q < 3, l < 72
constexpr size_t PSI_BYTES = sizeof(cx2) * P * S;     // 84096
constexpr size_t SMEM_BYTES = PSI_BYTES + sizeof(cpx) * 2 * P;

struct Wf {          // per-wavefield constants
    int py, px, lt, gx;
    bool x_in;
    const cpx* probe;
    cpx* ckpt;
    cpx* det;
};

ADM_DEV Wf make_wf(const MsParams& p, int wf, int l) {
    Wf w;
    const int b = wf / p.M, m = wf % p.M;
    w.py = p.pos[2 * b]; w.px = p.pos[2 * b + 1];
    w.lt = l + (w.px & 1);
    w.gx = w.px + l;
    w.x_in = (w.gx >= 0) && (w.gx < p.X);
    w.probe = p.probe + (size_t)(p.probe_per_pattern ? wf : m) * P * P;
    w.ckpt = p.ckpt + (size_t)wf * p.Nz * P * TP;
    w.det = p.det + (size_t)wf * P * P;
    return w;
}

// ---- packed passes over the paired wavefield (positions from Idx<Cfg72>, see fft_passes.cuh) ---------------------
template <bool XDIR> ADM_DEV int adr(int l, int along) { return XDIR ? l * S + along : along * S + l; }

template <bool XDIR, bool INV> ADM_DEV void pass_A(cx2* psi, const Idx<C>& I, int l) {
#pragma unroll
    for (int gi = 0; gi < C::NGA; ++gi) {
        cx2 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = psi[adr<XDIR>(l, I.a(j, gi))];
        dft8_2<INV>(v);
#pragma unroll
        for (int j = 0; j < 8; ++j) psi[adr<XDIR>(l, I.a(j, gi))] = v[j];
    }
}
// B pass along y; hy != null: multiply the outputs by the y factor of the transfer function (forward only)
template <bool INV> ADM_DEV void pass_B_y(cx2* psi, const cpx* hy, const Idx<C>& I, int l) {
#pragma unroll
    for (int gi = 0; gi < C::NGB; ++gi) {
        if (I.group(gi) < 8) {
            cx2 v[9];
#pragma unroll
            for (int n = 0; n < 9; ++n) v[n] = psi[I.b(n, gi) * S + l];
            dft9_2<INV>(v);
#pragma unroll
            for (int n = 0; n < 9; ++n) psi[I.b(n, gi) * S + l] = (hy != nullptr) ? cmul2s(v[n], hy[I.b(n, gi)]) : v[n];
        }
    }
}
// P4: B pass along x, transfer function (x factor, carries 1/P^2), inverse B pass
ADM_DEV void pass_BHB_x(cx2* psi, const cpx* hx, const Idx<C>& I, int l) {
#pragma unroll
    for (int gi = 0; gi < C::NGB; ++gi) {
        if (I.group(gi) < 8) {
            cx2 v[9];
#pragma unroll
            for (int n = 0; n < 9; ++n) v[n] = psi[l * S + I.b(n, gi)];
            dft9_2<false>(v);
#pragma unroll
            for (int n = 0; n < 9; ++n) v[n] = cmul2s(v[n], hx[I.b(n, gi)]);
            dft9_2<true>(v);
#pragma unroll
            for (int n = 0; n < 9; ++n) psi[l * S + I.b(n, gi)] = v[n];
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 2) k_ms_x2_fwd(MsParams p) {
    ADM_SMEM_DECL
    cx2* psi = (cx2*)smem_raw;
    cpx* hs = (cpx*)(smem_raw + PSI_BYTES);
    const int t = threadIdx.x;
    const bool active = t < C::NACT;
    const int q = t / P, l = t % P;
    const Idx<C> I(q);
    const int nwf = p.B * p.M;
    const int wfA = 2 * blockIdx.x;
    const bool validB = wfA + 1 < nwf;
    const Wf A = make_wf(p, wfA, l), Bw = make_wf(p, validB ? wfA + 1 : wfA, l);
    const size_t slice_stride = (size_t)p.Y * p.Xp;

    if (t < 2 * P) hs[t] = p.Hp[t];
    __syncthreads();

    for (int z = 0; z < p.Nz; ++z) {
        const bool last = (z == p.Nz - 1);
        const bool go_on = !last || p.free_prop != ADM_PROP_NONE;
        // ---- P1: [Ay^-1] . checkpoint . modulation . [Ay] ------------------------------------------------
        if (active) {
#pragma unroll
            for (int gi = 0; gi < C::NGA; ++gi) {
                cpx ta[8], tb[8];
                // T' of this slice for both wavefields (L2-resident volume; issued first, the butterfly hides the latency)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int y = I.a(j, gi);
                    const int gya = A.py + y, gyb = Bw.py + y;
                    const bool ia = A.x_in && gya >= 0 && gya < p.Y, ib = Bw.x_in && gyb >= 0 && gyb < p.Y;
                    ta[j] = ia ? p.tvol[(size_t)z * slice_stride + (size_t)gya * p.Xp + A.gx] : mk(0.f, 0.f);
                    tb[j] = ib ? p.tvol[(size_t)z * slice_stride + (size_t)gyb * p.Xp + Bw.gx] : mk(0.f, 0.f);
                }
                cx2 v[8];
                if (z == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = pack_cx(A.probe[I.a(j, gi) * P + l], Bw.probe[I.a(j, gi) * P + l]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = psi[I.a(j, gi) * S + l];
                    dft8_2<true>(v);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int y = I.a(j, gi);
                    cpx a, b;
                    unpack_cx(v[j], a, b);
                    A.ckpt[(size_t)z * P * TP + y * TP + A.lt] = a;
                    if (validB) Bw.ckpt[(size_t)z * P * TP + y * TP + Bw.lt] = b;
                    a = modulate(a, ta[j]);
                    b = modulate(b, tb[j]);
                    if (!go_on) {   // exit wave, natural order
                        A.det[y * P + l] = a;
                        if (validB) Bw.det[y * P + l] = b;
                    }
                    if (last && p.ff_conj) { a = cconj(a); b = cconj(b); }
                    v[j] = pack_cx(a, b);
                }
                if (go_on) {
                    dft8_2<false>(v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) psi[I.a(j, gi) * S + l] = v[j];
                }
            }
        }
        if (!go_on) break;
        __syncthreads();
        if (last) {
            // far field: finish the forward 2-D FFT (By, Ax, Bx) and store fftshift-ed, natural order
            if (active) pass_B_y<false>(psi, nullptr, I, l);
            __syncthreads();
            if (active) pass_A<true, false>(psi, I, l);
            __syncthreads();
            if (active) {
                const int fy = (perm_freq<C>(l) + P / 2) % P;
#pragma unroll
                for (int gi = 0; gi < C::NGB; ++gi) {
                    if (I.group(gi) < 8) {
                        cx2 v[9];
#pragma unroll
                        for (int n = 0; n < 9; ++n) v[n] = psi[l * S + I.b(n, gi)];
                        dft9_2<false>(v);
#pragma unroll
                        for (int k = 0; k < 9; ++k) {
                            const int fx = (perm_freq<C>(I.b(k, gi)) + P / 2) % P;
                            cpx a, b;
                            unpack_cx(v[k], a, b);
                            a = cscale(a, p.ff_scale); b = cscale(b, p.ff_scale);
                            if (p.ff_conj) { a = cconj(a); b = cconj(b); }
                            A.det[fy * P + fx] = a;
                            if (validB) Bw.det[fy * P + fx] = b;
                        }
                    }
                }
            }
            break;
        }
        // ---- P2 .. P6 --------------------------------------------------------------------------------------
        if (active) pass_B_y<false>(psi, hs, I, l);
        __syncthreads();
        if (active) pass_A<true, false>(psi, I, l);
        __syncthreads();
        if (active) pass_BHB_x(psi, hs + P, I, l);
        __syncthreads();
        if (active) pass_A<true, true>(psi, I, l);
        __syncthreads();
        if (active) pass_B_y<true>(psi, nullptr, I, l);
        __syncthreads();
    }
}
}   // namespace msx2

// ================================================================================================
// host side
// ================================================================================================
// Opt-in (ADM_MS_X2=1).  Measured on B200 at config D (profiles/r01g_*): 4.26 ms against 4.04 ms for the one-wavefield
// kernel -- the packed butterflies halve the FP instructions (1.79 G -> 0.85 G warp instructions) but pairing two patches
// with different origins adds per-point address arithmetic, bounds predicates and pack / unpack moves in P1 (MOV 15 %
// of the mix), and dependent FADD2 / FFMA2 chains wait twice as long ('wait' 20 % of the stall samples).
int adm_msx2_fwd_supported(const adm_multislice_cfg* c) {
    const char* e = getenv("ADM_MS_X2");
    if (!e || !atoi(e)) return 0;
    return c->P == 72 && c->h_separable && c->n_slices > 1 && c->free_prop != ADM_PROP_NEAR && c->B * c->M >= 2;
}

int adm_msx2_fwd_launch(const adm_multislice_cfg* c, const MsParams& p, void* stream) {
    const int nwf = c->B * c->M;
    if (set_smem(msx2::k_ms_x2_fwd, msx2::SMEM_BYTES)) return ADM_ERR_CUDA;
    ADM_LAUNCH(msx2::k_ms_x2_fwd, (nwf + 1) / 2, msx2::NTHREADS, msx2::SMEM_BYTES, stream, p);
    return ADM_OK;
}
