// Three-pass multislice kernels for 72- and 64-pixel probes and a separable (paraxial) transfer function  (sm_100a).
// Included by multislice.cu (after MsSmem / the balanced-schedule helpers); same C ABI, same checkpoint / tile /
// hand-off formats as the five-pass kernels there, which remain the path for single-slice objects, a general
// (non-paraxial) transfer function and the adjoint-wavefield checkpoints (adm_multislice_bwd_gpsi).
// Everything below is written for P = RA x RB per axis: 72 = 3 x 24 (described here) and 64 = 4 x 16.
//
// Why: the five-pass kernels sit on the shared-memory pipe (ncu r02y: l1tex 74 %, ~4.8 k wavefronts per slice, of which
// 5 x 648 are the five round trips of the wavefield).  Here 72 = 3 x 24 per axis (Cooley-Tukey, twiddle W72^(r k1) between
// the stages):  n = 24 j + r  (j = 0..2, r = 0..23);  stage A = radix 3 over j, stage B = radix 24 over r (in registers as
// a twiddle-free 3 x 8 prime-factor transform).  The radix-3 stages of BOTH axes, their inverses and the point-wise
// slice modulation act on the same 3 x 3 points {(24 jy + ry, 24 jx + rx)}, so a slice is THREE round trips:
//
//   M   [Ax^-1 Ay^-1 . checkpoint . psi (1 + T') . Ay Ax]     one thread per (ry, rx): 9 values, 576 blocks
//   By  [By . hy . By^-1]                                      one thread per (k1y, column): 24 values, 216 groups
//   Bx  [Bx . hx . Bx^-1]                                      one thread per (k1x, row):    24 values, 216 groups
//
// and three CTA barriers.  The register blocks stay small (9 / 24 values; 216 threads as before, every thread busy in
// the B passes), the butterfly count rises by ~13 % (twiddles, radix 24 instead of 8 + 9), the shared-memory
// wavefronts fall by ~30 %.
//
// Index conventions along one axis: after stage A the value (k1, r) sits at position 24 k1 + r; after the forward
// 24-point transform register slot s of group k1 holds the frequency k1 + 3 f24(s); the transfer-function tables are
// laid out in that order (ms3_hs_fill), and the inverse transform takes slot order back to natural order, so no reorder
// pass exists.  Overall scale 1/72^2 is carried by the hx table (adm_prepare_transfer_function_separable).
//
// Reference behaviour: adorym/propagate.py:208-280 (slice loop, detector transform), as multislice.cu.
#pragma once
#include <type_traits>

// P = RA * RB per axis (72 = 3 x 24, 64 = 4 x 16).  NT threads per CTA; the B passes use the first P * RA of them (one
// RB-point group each), the M pass MT of them with MR blocks each (thread t owns the blocks (t / RB + (MT / RB) k, t % RB),
// k < MR, while t + MT k < RB^2).  MINB: CTAs per SM the registers are capped for.
template <int P_, int RA_, int NT_, int MT_, int MR_, int MINB_, bool COMB_ = false> struct Ms3Cfg {
    static constexpr int P = P_, RA = RA_, RB = P_ / RA_, S = P_ + 1, TP = P_ + 2, NACT = P_ * RA_, NTHREADS = NT_, MT = MT_, MR = MR_;
    static constexpr int RS = MT_ / RB, NB = RB * RB, MINB = MINB_;
    static constexpr bool COMB = COMB_;      // products of the y and x twiddles as per-thread constants (one multiply instead of two)
    static constexpr size_t TILE = sizeof(cpx) * P * TP;                 // T' tile (forward) / checkpoint + dL/dT tile (adjoint)
    static constexpr size_t PSI = sizeof(cpx) * P * S;
    static constexpr size_t HS = sizeof(cpx) * 2 * P;
    static constexpr size_t BYTES = TILE + PSI + HS + 16;
    static_assert((P_ == 72 && RA_ == 3) || (P_ == 64 && RA_ == 4), "72 = 3 x 24 or 64 = 4 x 16");
    static_assert(MT_ % RB == 0 && MT_ * MR_ >= NB && MT_ <= NT_ && NT_ >= NACT, "M-pass blocks / B-pass groups must be covered");
    static_assert(TILE % 128 == 0 && PSI % 16 == 0, "TMA / float4 alignment of the carve-up");
    // block round k exists for every thread
    ADM_HDM static constexpr bool full_round(int k) { return MT_ * (k + 1) <= NB; }
    // frequency (mod RB) held by register slot s after the forward RB-point transform
    ADM_HDM static constexpr int fB(int s) {
        return RB == 24 ? (16 * ((2 * s) % 3) + 9 * ((3 * s) % 8)) % 24 : s / 4 + 4 * (s % 4);
    }
    // position of frequency f in the permuted order of the five-pass kernels' separable tables
    // (adm_prepare_transfer_function_separable: entry p holds frequency perm_freq<Cfg72 / Cfg64>(p))
    ADM_HDM static constexpr int five_pass_pos(int f) {
        return P_ == 72 ? (9 * (f % 8) + 64 * ((9 - f % 9) % 9)) % 72 : 8 * (f % 8) + f / 8;
    }
};
typedef Ms3Cfg<72, 3, 224, 216, 3, 2> Ms3;      // adjoint: 7 warps, 128 registers (the T' values held across a barrier need them)
typedef Ms3Cfg<72, 3, 288, 288, 2, 2, true> Ms3w;     // forward, many wavefields: 9 warps, M pass in two balanced rounds, combined twiddles (96 registers: 2.66 ms)
typedef Ms3Cfg<64, 4, 256, 256, 1, 2> Ms3p64;   // 64-pixel probes: one 4 x 4 block and one 16-point group per thread

// 24-point DFT in registers: Good-Thomas 3 x 8, n = (8 a + 3 b) mod 24, k = (16 ka + 9 kb) mod 24, slot (8 ka + 3 kb) mod 24.
// INV = false: natural order in, slot order out.  INV = true: slot order in, natural order out (conjugate, unnormalised).
template <bool INV, class T> ADM_DEV void ms3_dft24(T* u) {
    if constexpr (!INV) {
#pragma unroll
        for (int b = 0; b < 8; ++b) dft3<false>(u[(3 * b) % 24], u[(8 + 3 * b) % 24], u[(16 + 3 * b) % 24]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        T v[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) v[b] = u[(8 * a + 3 * b) % 24];
        dft8<INV>(v);
#pragma unroll
        for (int b = 0; b < 8; ++b) u[(8 * a + 3 * b) % 24] = v[b];
    }
    if constexpr (INV) {
#pragma unroll
        for (int b = 0; b < 8; ++b) dft3<true>(u[(3 * b) % 24], u[(8 + 3 * b) % 24], u[(16 + 3 * b) % 24]);
    }
}

// 16-point DFT in registers: Cooley-Tukey 4 x 4, n = 4 n1 + n2, k = k1 + 4 k2, slot 4 k1 + k2; twiddles W16^(n2 k1).
// INV = false: natural order in, slot order out.  INV = true: slot order in, natural order out (conjugate, unnormalised).
#define ADM_C16_1 0.92387953251128675613f
#define ADM_S16_1 0.38268343236508977173f
template <bool INV, class T> ADM_DEV void ms3_tw16(T* a) {
    // a[4 k1 + n2] *= W16^(n2 k1) (forward) or its conjugate: exponents 1 2 3 / 2 4 6 / 3 6 9
    a[5] = tw9<INV>(a[5], ADM_C16_1, ADM_S16_1);
    a[6] = tw9<INV>(a[6], ADM_SQRT1_2, ADM_SQRT1_2);
    a[7] = tw9<INV>(a[7], ADM_S16_1, ADM_C16_1);
    a[9] = tw9<INV>(a[9], ADM_SQRT1_2, ADM_SQRT1_2);
    a[10] = tw9<INV>(a[10], 0.0f, 1.0f);
    a[11] = tw9<INV>(a[11], -ADM_SQRT1_2, ADM_SQRT1_2);
    a[13] = tw9<INV>(a[13], ADM_S16_1, ADM_C16_1);
    a[14] = tw9<INV>(a[14], -ADM_SQRT1_2, ADM_SQRT1_2);
    a[15] = tw9<INV>(a[15], -ADM_C16_1, -ADM_S16_1);
}
template <bool INV, class T> ADM_DEV void ms3_dft16(T* a) {
    if constexpr (!INV) {
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) dft4<false>(a[n2], a[4 + n2], a[8 + n2], a[12 + n2]);
        ms3_tw16<false>(a);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) dft4<false>(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
    } else {
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) dft4<true>(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
        ms3_tw16<true>(a);
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) dft4<true>(a[n2], a[4 + n2], a[8 + n2], a[12 + n2]);
    }
}
template <int RB, bool INV, class T> ADM_DEV void ms3_dftB(T* u) {
    if constexpr (RB == 24) ms3_dft24<INV>(u);
    else ms3_dft16<INV>(u);
}

// W_P^n = exp(-2 pi i n / P)
template <class T, int P> ADM_DEV T ms3_wP(int n) {
#ifndef ADM_HOST_EMU
    float s, c;
    sincospif(2.0f * (float)n / (float)P, &s, &c);
    return tmk<T>(c, -s);
#else
    double a = 2.0 * 3.14159265358979323846 * n / P;
    return tmk<T>((float)cos(a), (float)-sin(a));
#endif
}

// transfer-function tables in slot order: hs[dir][k1][s] = h_dir[k1 + RA fB(s)].  Hsep = [hy | hx] in the permuted order of
// the five-pass kernels.
template <class C, class T> ADM_DEV void ms3_hs_fill(T* hs, const T* Hsep, int t) {
    if (t < 2 * C::P) {
        const int dir = t / C::P, i = t % C::P;
        const int f = i / C::RB + C::RA * C::fB(i % C::RB);
        hs[t] = Hsep[dir * C::P + C::five_pass_pos(f)];
    }
}

// The same tables for the free-space (near-field) transfer function, taken out of the prepared GENERAL table
// Hf[px][py] = scale H[freq(py)][freq(px)] (adm_prepare_transfer_function; positions in the five-pass order, position 0
// holds frequency 0) of an outer product H = hy (x) hx:  hy'(f) = Hf[0][pos(f)] = scale H(f, 0),
// hx'(f) = Hf[pos(f)][0] / Hf[0][0] = H(0, f) / H(0, 0).
template <class C, class T> ADM_DEV void ms3_hf_fill(T* hs, const T* Hf, int t) {
    if (t < 2 * C::P) {
        const int dir = t / C::P, i = t % C::P;
        const int pos = C::five_pass_pos(i / C::RB + C::RA * C::fB(i % C::RB));
        if (dir == 0) hs[t] = Hf[pos];
        else {
            const T a = Hf[pos * C::P], b = Hf[0];
            const float n = 1.0f / (cx(b) * cx(b) + cy(b) * cy(b));
            hs[t] = cscale(cmulc(a, b), n);
        }
    }
}

// radix-RA stages of both axes on an RA x RA block w[jy][jx] (+ the twiddles W_P^(ry k1y) W_P^(rx k1x)); wy / wx = {W^r, W^2r, ..}
template <bool INV, int RA, class T> ADM_DEV void ms3_dftA(T& a0, T& a1, T& a2, T& a3) {
    if constexpr (RA == 3) dft3<INV>(a0, a1, a2);
    else dft4<INV>(a0, a1, a2, a3);
}
template <bool INV, int RA, bool COMB, class T> ADM_DEV void ms3_A(T (&w)[RA][RA], const T* wy, const T* wx, const T* wyx) {
    // twiddle of element (k1y, k1x): wy[k1y-1] wx[k1x-1]; COMB: wyx holds the products, row-major (k1y-1, k1x-1), one multiply
    // per element (needs (RA-1)^2 more constants per block; measured: forward 2.71 -> 2.66 ms, adjoint slower -- registers)
    auto twiddle = [&](auto conj_tag) {
        constexpr bool CJ = decltype(conj_tag)::value;
        if constexpr (COMB) {
#pragma unroll
            for (int ky = 0; ky < RA; ++ky)
#pragma unroll
                for (int kx = 0; kx < RA; ++kx) {
                    if (ky == 0 && kx == 0) continue;
                    const T f = (ky > 0 && kx > 0) ? wyx[(ky - 1) * (RA - 1) + kx - 1] : (ky > 0 ? wy[ky - 1] : wx[kx - 1]);
                    w[ky][kx] = CJ ? cmulc(w[ky][kx], f) : cmul(w[ky][kx], f);
                }
        } else {
#pragma unroll
            for (int k = 1; k < RA; ++k)
#pragma unroll
                for (int j = 0; j < RA; ++j) w[k][j] = CJ ? cmulc(w[k][j], wy[k - 1]) : cmul(w[k][j], wy[k - 1]);
#pragma unroll
            for (int k = 1; k < RA; ++k)
#pragma unroll
                for (int j = 0; j < RA; ++j) w[j][k] = CJ ? cmulc(w[j][k], wx[k - 1]) : cmul(w[j][k], wx[k - 1]);
        }
    };
    if constexpr (INV) twiddle(std::true_type());
#pragma unroll
    for (int j = 0; j < RA; ++j) ms3_dftA<INV, RA>(w[j][0], w[j][1], w[j][2], w[j][RA - 1]);
#pragma unroll
    for (int j = 0; j < RA; ++j) ms3_dftA<INV, RA>(w[0][j], w[1][j], w[2][j], w[RA - 1][j]);
    if constexpr (!INV) twiddle(std::false_type());
}

// per-thread constants of the M pass: thread t owns the blocks (ry0 + RS k, rx), k < MR
template <class T, class C> struct Ms3Thread {
    int ry0, rx;
    T wy[C::MR][C::RA - 1], wx[C::RA - 1];
    T wyx[C::COMB ? C::MR : 1][(C::RA - 1) * (C::RA - 1)];
    ADM_HDM const T* comb(int k) const { return C::COMB ? wyx[k] : nullptr; }
};
template <class T, class C> ADM_DEV void ms3_thread_init(Ms3Thread<T, C>& I, int t) {
    I.ry0 = (t / C::RB) % C::RB; I.rx = t % C::RB;
#pragma unroll
    for (int k = 0; k < C::MR; ++k)
#pragma unroll
        for (int e = 1; e < C::RA; ++e) I.wy[k][e - 1] = ms3_wP<T, C::P>(e * ((I.ry0 + C::RS * k) % C::RB));
#pragma unroll
    for (int e = 1; e < C::RA; ++e) I.wx[e - 1] = ms3_wP<T, C::P>(e * I.rx);
    if constexpr (C::COMB) {
#pragma unroll
        for (int k = 0; k < C::MR; ++k)
#pragma unroll
            for (int ey = 1; ey < C::RA; ++ey)
#pragma unroll
                for (int ex = 1; ex < C::RA; ++ex)
                    I.wyx[k][(ey - 1) * (C::RA - 1) + ex - 1] = ms3_wP<T, C::P>((ey * ((I.ry0 + C::RS * k) % C::RB) + ex * I.rx) % C::P);
    }
}

// B pass along y (YDIR, lanes along x: l = column) or along x (lanes along y: l = row) for group q = k1.
// MODE 0: forward . h . inverse;  1: forward only (slot order left in place);  2: inverse only (slot order in).
template <class C, bool YDIR, int MODE, class T> ADM_DEV void ms3_B(T* psi, const T* hsd, int q, int l) {
    constexpr int S = C::S, RB = C::RB;
    T* base = psi + (YDIR ? RB * q * S + l : l * S + RB * q);
    constexpr int st = YDIR ? S : 1;
    T u[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) u[r] = base[r * st];
    if constexpr (MODE != 2) ms3_dftB<RB, false>(u);
    if constexpr (MODE == 0) {
        const T* hq = hsd + RB * q;
#pragma unroll
        for (int s = 0; s < RB; s += 2) {
            T ha, hb;
            lds_pair(hq + s, ha, hb);
            u[s] = cmul(u[s], ha);
            u[s + 1] = cmul(u[s + 1], hb);
        }
    }
    if constexpr (MODE != 1) ms3_dftB<RB, true>(u);
#pragma unroll
    for (int r = 0; r < RB; ++r) base[r * st] = u[r];
}

// ================================================================================================
// Forward
// ================================================================================================
template <class T, class C> __global__ void __launch_bounds__(C::NTHREADS, C::MINB) k_ms3_fwd(const ADM_GRID_CONSTANT AdmTileMap tmap, MsParams p) {
    ADM_SMEM_DECL
    constexpr int P = C::P, S = C::S, TP = C::TP, RA = C::RA, RB = C::RB;
    constexpr unsigned TILE_BYTES = (unsigned)C::TILE;
    T* tile = (T*)smem_raw;
    T* psi = tile + P * TP;
    T* hs = psi + P * S;
    adm_mbar_t* bar = (adm_mbar_t*)(hs + 2 * P);
    const int t = threadIdx.x;
    const bool active = t < C::NACT;          // B passes
    const bool mactive = t < C::MT;           // M pass
    const int q = t / P, l = t % P;
    Ms3Thread<T, C> I;
    ms3_thread_init<T, C>(I, t);
    const int c = blockIdx.x;

    ms3_hs_fill<C>(hs, (const T*)p.Hp, t);
    if (t == 0) mbar_init(bar, 1);
    unsigned ph = 0;

    int wf_first, wf_last;
    long long s0, s1;
    ms_nseg(p, c, wf_first, wf_last, s0, s1);
    for (int wfi = wf_last; wfi >= wf_first; --wfi) {
        const MsSeg sg = ms_seg(p, wfi, s0, s1);
        const int wf = sg.wf;
        const int b = wf / p.M, m = wf % p.M;
        const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
        const int pxa = px & ~1, xo = I.rx + (px & 1);      // even TMA start column, this thread's first column inside the tile
        const T* probe = (const T*)p.probe + (size_t)(p.probe_per_pattern ? wf : m) * P * P;
        T* ckpt = (T*)p.ckpt + (size_t)wf * p.Nz * P * TP;
        T* det = (T*)p.det + (size_t)wf * P * P;

        __syncthreads();                    // the previous segment is done with the tile and the wavefield
        auto issue_tile = [&](int zz) {      // one elected thread
            mbar_expect_tx(bar, TILE_BYTES);
            tma_load_tile(tile, &tmap, 2 * pxa, py, zz, bar);
        };
        if (t == 0) issue_tile(sg.zb);
        if (sg.zb > 0) {                    // late part of a cut chain: the wavefield comes from CTA c - 1
            if (t == 0) flag_wait_and_clear(p.flags + c);
            __syncthreads();
            for (int i = t; i < (int)(C::PSI / 16); i += C::NTHREADS) {
#ifndef ADM_HOST_EMU
                ((float4*)psi)[i] = __ldcg(p.handoff + (size_t)c * p.slot_f4 + i);
#else
                ((float4*)psi)[i] = p.handoff[(size_t)c * p.slot_f4 + i];
#endif
            }
            __syncthreads();
        }

        if (sg.zb == 0 && mactive) {
            // fresh wavefield: the probe enters the loop in the state every slice leaves behind (stage A applied along both
            // axes, divided by the 9 that the unnormalised inverse stages return), so that the M pass has ONE code path.
            // Every thread reads back exactly the positions it writes here: no barrier.
#pragma unroll
            for (int k = 0; k < C::MR; ++k) {
                if (!C::full_round(k) && t + C::MT * k >= C::NB) break;
                const int ry = I.ry0 + C::RS * k;
                T w[RA][RA];
#pragma unroll
                for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                    for (int jx = 0; jx < RA; ++jx) w[jy][jx] = cscale(probe[(RB * jy + ry) * P + RB * jx + I.rx], 1.0f / (float)(RA * RA));
                ms3_A<false, RA, C::COMB>(w, I.wy[k], I.wx, I.comb(k));
#pragma unroll
                for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                    for (int jx = 0; jx < RA; ++jx) psi[(RB * jy + ry) * S + RB * jx + I.rx] = w[jy][jx];
            }
        }

        for (int z = sg.zb; z < sg.ze; ++z) {
            const bool last = (z == p.Nz - 1);
            const bool go_on = !last || p.free_prop != ADM_PROP_NONE;
            // per-gap transfer functions (sparse multislice): the previous slice's last reader (its Bx pass) is one barrier
            // back, the first reader (By) comes after the barrier that closes M
            if (!last && p.h_per_slice) ms3_hs_fill<C>(hs, (const T*)p.Hp + (size_t)z * 2 * P, t);
            // ---- M: [Ax^-1 Ay^-1] . checkpoint . modulation . [Ay Ax] ------------------------------------------
            // (ONE code path for all slices with run-time checks of the last-slice options.  Measured alternatives: a second,
            // branch-free instantiation for the other slices inside the loop 3.5 % slower; the last slice as a second
            // instantiation BEHIND the loop 3 % slower; the detector transforms inside the loop 2 % slower.)
            if (mactive) {
                mbar_wait(bar, ph);
                T* ck = ckpt + (size_t)z * P * TP;
#pragma unroll
                for (int k = 0; k < C::MR; ++k) {
                    if (!C::full_round(k) && t + C::MT * k >= C::NB) break;
                    const int ry = I.ry0 + C::RS * k;
                    T w[RA][RA];
#pragma unroll
                    for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                        for (int jx = 0; jx < RA; ++jx) w[jy][jx] = psi[(RB * jy + ry) * S + RB * jx + I.rx];
                    ms3_A<true, RA, C::COMB>(w, I.wy[k], I.wx, I.comb(k));
#pragma unroll
                    for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                        for (int jx = 0; jx < RA; ++jx) {
                            const int o = (RB * jy + ry) * TP + RB * jx + xo;
                            ck[o] = w[jy][jx];
                            w[jy][jx] = modulate(w[jy][jx], tile[o]);
                        }
                    if (!go_on) {
#pragma unroll
                        for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                            for (int jx = 0; jx < RA; ++jx) det[(RB * jy + ry) * P + RB * jx + I.rx] = w[jy][jx];   // exit wave
                        continue;
                    }
                    if (last && p.ff_conj) {
#pragma unroll
                        for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                            for (int jx = 0; jx < RA; ++jx) w[jy][jx] = cconj(w[jy][jx]);
                    }
                    ms3_A<false, RA, C::COMB>(w, I.wy[k], I.wx, I.comb(k));
#pragma unroll
                    for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                        for (int jx = 0; jx < RA; ++jx) psi[(RB * jy + ry) * S + RB * jx + I.rx] = w[jy][jx];
                }
            }
            ph ^= 1;
            __syncthreads();
            if (t == 0 && z + 1 < sg.ze) issue_tile(z + 1);   // every thread has consumed tile(z): prefetch tile(z+1) behind the B passes
            if (last) break;      // exit wave: stored by the M pass; far / near field: detector transform behind the loop
            if (active) ms3_B<C, true, 0>(psi, hs, q, l);
            __syncthreads();
            if (active) ms3_B<C, false, 0>(psi, hs + P, q, l);
            __syncthreads();
        }
        // ---- detector transform of a chain that reached the last slice (kept out of the slice loop: code size) ----
        if (sg.ze == p.Nz && p.free_prop == ADM_PROP_FAR) {
            // far field: finish the forward 2-D transform and store fftshift-ed, natural order
            if (active) ms3_B<C, true, 1>(psi, hs, q, l);
            __syncthreads();
            if (active) {
                T u[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) u[r] = psi[l * S + RB * q + r];
                ms3_dftB<RB, false>(u);
                const int fy = (l / RB + RA * C::fB(l % RB) + P / 2) % P;
#pragma unroll
                for (int s = 0; s < RB; ++s) {
                    const int fx = (q + RA * C::fB(s) + P / 2) % P;
                    T o = cscale(u[s], p.ff_scale);
                    det[fy * P + fx] = p.ff_conj ? cconj(o) : o;
                }
            }
        } else if (sg.ze == p.Nz && p.free_prop == ADM_PROP_NEAR) {
            // near field: the two B passes with the free-space tables, closing inverse stages, natural order store; then the
            // in-object tables again (their next reader is behind the barriers of the next wavefield)
            ms3_hf_fill<C>(hs, (const T*)p.Hf, t);
            __syncthreads();
            if (active) ms3_B<C, true, 0>(psi, hs, q, l);
            __syncthreads();
            if (active) ms3_B<C, false, 0>(psi, hs + P, q, l);
            __syncthreads();
            if (mactive) {
#pragma unroll
                for (int k = 0; k < C::MR; ++k) {
                    if (!C::full_round(k) && t + C::MT * k >= C::NB) break;
                    const int ry = I.ry0 + C::RS * k;
                    T w[RA][RA];
#pragma unroll
                    for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                        for (int jx = 0; jx < RA; ++jx) w[jy][jx] = psi[(RB * jy + ry) * S + RB * jx + I.rx];
                    ms3_A<true, RA, C::COMB>(w, I.wy[k], I.wx, I.comb(k));
#pragma unroll
                    for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                        for (int jx = 0; jx < RA; ++jx) det[(RB * jy + ry) * P + RB * jx + I.rx] = w[jy][jx];
                }
            }
            if (!p.h_per_slice) ms3_hs_fill<C>(hs, (const T*)p.Hp, t);
        }
        if (sg.ze < p.Nz) {                 // early part of a cut chain: hand the wavefield to CTA c + 1
            float4* dst = p.handoff + (size_t)(c + 1) * p.slot_f4;
            for (int i = t; i < (int)(C::PSI / 16); i += C::NTHREADS) dst[i] = ((const float4*)psi)[i];
#ifndef ADM_HOST_EMU
            __threadfence();
#endif
            __syncthreads();
            if (t == 0) flag_set(p.flags + c + 1);
        }
    }
}

// ================================================================================================
// Adjoint
// ================================================================================================
template <class T, class C> __global__ void __launch_bounds__(C::NTHREADS, C::MINB) k_ms3_bwd(const ADM_GRID_CONSTANT AdmTileMap gmap, MsParams p) {
    ADM_SMEM_DECL
    constexpr int P = C::P, S = C::S, TP = C::TP, RA = C::RA, RB = C::RB;
    constexpr unsigned TILE_BYTES = (unsigned)C::TILE;
    T* ckb = (T*)smem_raw;          // psi checkpoint of the slice (bulk copy); overwritten in place by dL/dT
    T* psi = ckb + P * TP;          // running adjoint wavefield
    T* hs = psi + P * S;
    adm_mbar_t* bar = (adm_mbar_t*)(hs + 2 * P);
    const int t = threadIdx.x;
    const bool active = t < C::NACT;          // B passes
    const bool mactive = t < C::MT;           // M pass
    const int q = t / P, l = t % P;
    Ms3Thread<T, C> I;
    ms3_thread_init<T, C>(I, t);
    const int c = gridDim.x - 1 - blockIdx.x;

    ms3_hs_fill<C>(hs, (const T*)p.Hp, t);
    if (t == 0) mbar_init(bar, 1);
    unsigned phase = 0;

    int wf_first, wf_last;
    long long s0, s1;
    ms_nseg(p, c, wf_first, wf_last, s0, s1);
    for (int wfi = wf_first; wfi <= wf_last; ++wfi) {
        const MsSeg sg = ms_seg(p, wfi, s0, s1);
        const int wf = sg.wf;
        const int b = wf / p.M, m = wf % p.M;
        const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
        const int pxa = px & ~1, xo = I.rx + (px & 1);
        const T* ckpt = (T*)p.ckpt + (size_t)wf * p.Nz * P * TP;
        const T* gdet = (const T*)p.det + (size_t)wf * P * P;
        const size_t slice_stride = (size_t)p.Y * p.Xp;
        // T' of one slice for this thread's blocks: plain loads of the L2-resident volume, ISSUED one barrier early (behind
        // the last pass of the previous slice) so that their latency overlaps the barrier wait
        T tpa[C::MR][RA][RA];
        auto load_tp = [&](int zz) {
            const T* tzz = (const T*)p.tvol + (size_t)zz * slice_stride;
#pragma unroll
            for (int k = 0; k < C::MR; ++k) {
                if (!C::full_round(k) && t + C::MT * k >= C::NB) break;
#pragma unroll
                for (int jy = 0; jy < RA; ++jy) {
                    const int gy = py + RB * jy + I.ry0 + C::RS * k;
#pragma unroll
                    for (int jx = 0; jx < RA; ++jx) {
                        const int gx = px + RB * jx + I.rx;
                        const bool inb = (unsigned)gy < (unsigned)p.Y && (unsigned)gx < (unsigned)p.X;
                        tpa[k][jy][jx] = inb ? tzz[(unsigned)(gy * p.Xp + gx)] : tmk<T>(0.f, 0.f);
                    }
                }
            }
        };
        if (mactive) load_tp(sg.ze - 1);

        __syncthreads();                    // the previous segment is done with the wavefield
        if (t == 0) {
            bulk_wait_read_all();           // ... and its last reduce has drained the tile
            mbar_expect_tx(bar, TILE_BYTES);
            bulk_load(ckb, ckpt + (size_t)(sg.ze - 1) * P * TP, TILE_BYTES, bar);
        }

        if (sg.ze < p.Nz) {
            if (t == 0) flag_wait_and_clear(p.flags + c + 1);
            __syncthreads();
            for (int i = t; i < (int)(C::PSI / 16); i += C::NTHREADS) {
#ifndef ADM_HOST_EMU
                ((float4*)psi)[i] = __ldcg(p.handoff + (size_t)(c + 1) * p.slot_f4 + i);
#else
                ((float4*)psi)[i] = p.handoff[(size_t)(c + 1) * p.slot_f4 + i];
#endif
            }
            __syncthreads();
        } else if (p.free_prop == ADM_PROP_FAR) {
            // adjoint of the detector transform: slot order in, inverse B along x, inverse B along y
            if (active) {
                T u[RB];
                const int fy = (l / RB + RA * C::fB(l % RB) + P / 2) % P;
#pragma unroll
                for (int s = 0; s < RB; ++s) {
                    const int fx = (q + RA * C::fB(s) + P / 2) % P;
                    T o = cscale(gdet[fy * P + fx], p.ff_scale);
                    u[s] = p.ff_conj ? cconj(o) : o;
                }
                ms3_dftB<RB, true>(u);
#pragma unroll
                for (int r = 0; r < RB; ++r) psi[l * S + RB * q + r] = u[r];
            }
            __syncthreads();
            if (active) ms3_B<C, true, 2>(psi, hs, q, l);
            __syncthreads();
        } else if (mactive) {
            // exit-wave / near-field detection: dL/d(detector wave) enters in the state every slice leaves behind (as the probe
            // does in the forward kernel); every thread reads back exactly the positions it writes here: no barrier.  (The
            // state carries 1 / RA^2, which the unnormalised inverse stages of M' return; the B passes of the near-field
            // adjoint below bring that factor themselves.)
            const float pre = p.free_prop == ADM_PROP_NEAR ? 1.0f : 1.0f / (float)(RA * RA);
#pragma unroll
            for (int k = 0; k < C::MR; ++k) {
                if (!C::full_round(k) && t + C::MT * k >= C::NB) break;
                const int ry = I.ry0 + C::RS * k;
                T w[RA][RA];
#pragma unroll
                for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                    for (int jx = 0; jx < RA; ++jx) w[jy][jx] = cscale(gdet[(RB * jy + ry) * P + RB * jx + I.rx], pre);
                ms3_A<false, RA, C::COMB>(w, I.wy[k], I.wx, I.comb(k));
#pragma unroll
                for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                    for (int jx = 0; jx < RA; ++jx) psi[(RB * jy + ry) * S + RB * jx + I.rx] = w[jy][jx];
            }
        }

        if (sg.ze == p.Nz && p.free_prop == ADM_PROP_NEAR) {
            // adjoint of the free-space propagation to the detector: the two B passes with the conjugate free-space tables
            ms3_hf_fill<C>(hs, (const T*)p.Hf, t);
            __syncthreads();
            if (active) ms3_B<C, true, 0>(psi, hs, q, l);
            __syncthreads();
            if (active) ms3_B<C, false, 0>(psi, hs + P, q, l);
            __syncthreads();
            if (!p.h_per_slice) ms3_hs_fill<C>(hs, (const T*)p.Hp, t);      // next reader: behind the barrier that closes M'
        }

        for (int z = sg.ze - 1; z >= sg.zb; --z, phase ^= 1) {
            const bool last = (z == p.Nz - 1);
            if (z > 0 && p.h_per_slice) ms3_hs_fill<C>(hs, (const T*)p.Hp + (size_t)(z - 1) * 2 * P, t);
            // ---- M': [Ax^-1 Ay^-1] . dL/dT = g' conj(psi_z) . g = g' conj(1 + T') . [Ay Ax] ----------------------
            // (three instantiations: KIND 1 = last slice (detector options), 2 = slice 0 (probe gradient, nothing to
            // propagate), 0 = every other slice, branch free; Nz > 1 on this path, so 1 and 2 never coincide.  Measured 0.5 %
            // faster than one code path with run-time checks -- the opposite of the forward kernel)
            auto m_pass = [&](auto kind_tag) {
                constexpr int KIND = decltype(kind_tag)::value;
#pragma unroll
                for (int k = 0; k < C::MR; ++k) {
                    if (!C::full_round(k) && t + C::MT * k >= C::NB) break;
                    const int ry = I.ry0 + C::RS * k;
                    T w[RA][RA];
#pragma unroll
                    for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                        for (int jx = 0; jx < RA; ++jx) w[jy][jx] = psi[(RB * jy + ry) * S + RB * jx + I.rx];
                    ms3_A<true, RA, C::COMB>(w, I.wy[k], I.wx, I.comb(k));
                    if (KIND == 1 && p.ff_conj) {
#pragma unroll
                        for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                            for (int jx = 0; jx < RA; ++jx) w[jy][jx] = cconj(w[jy][jx]);
                    }
                    if (k == 0) {
                        mbar_wait(bar, phase);
                        // the two pad columns of the tile are never written by the forward pass: add zeros there
                        if (t < 2 * P) ckb[(t >> 1) * TP + ((t & 1) ? P + 1 : ((px & 1) ? 0 : P))] = tmk<T>(0.f, 0.f);
                    }
#pragma unroll
                    for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                        for (int jx = 0; jx < RA; ++jx) {
                            const int o = (RB * jy + ry) * TP + RB * jx + xo;
                            const T wz = ckb[o];                            // psi_z (wave entering the slice)
                            ckb[o] = cmulc(w[jy][jx], wz);                  // dL/dT = g' conj(psi_z)
                            w[jy][jx] = modulate_conj(w[jy][jx], tpa[k][jy][jx]);
                        }
                    if constexpr (KIND != 2) {
                        ms3_A<false, RA, C::COMB>(w, I.wy[k], I.wx, I.comb(k));
#pragma unroll
                        for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                            for (int jx = 0; jx < RA; ++jx) psi[(RB * jy + ry) * S + RB * jx + I.rx] = w[jy][jx];
                    } else if (p.g_probe != nullptr) {
#pragma unroll
                        for (int jy = 0; jy < RA; ++jy)
#pragma unroll
                            for (int jx = 0; jx < RA; ++jx) {
                                const int o = (RB * jy + ry) * P + RB * jx + I.rx;
                                if (p.probe_per_pattern) ((T*)p.g_probe)[(size_t)wf * P * P + o] = w[jy][jx];
                                else red_add2(&p.g_probe[(size_t)m * P * P + o], cx(w[jy][jx]), cy(w[jy][jx]));
                            }
                    }
                }
                fence_proxy_async();      // the dL/dT tile is about to be read by the TMA unit
            };
            if (mactive) {
                if (z == 0) m_pass(std::integral_constant<int, 2>());
                else if (last) m_pass(std::integral_constant<int, 1>());
                else m_pass(std::integral_constant<int, 0>());
            }
            __syncthreads();
            if (t == 0 && pxa + P >= 0 && py + P >= 0) {
                tma_reduce_add_tile(&gmap, 2 * (pxa + P), py + P, z, ckb);
                bulk_commit();
            }
            if (z == 0) { phase ^= 1; break; }
            if (active) ms3_B<C, true, 0>(psi, hs, q, l);
            __syncthreads();
            if (t == 0 && z > sg.zb) {   // the reduce has drained the tile: prefetch the checkpoint of slice z-1 behind the x pass
                bulk_wait_read_all();
                mbar_expect_tx(bar, TILE_BYTES);
                bulk_load(ckb, ckpt + (size_t)(z - 1) * P * TP, TILE_BYTES, bar);
            }
            if (active) ms3_B<C, false, 0>(psi, hs + P, q, l);
            if (mactive && z > sg.zb) load_tp(z - 1);
            __syncthreads();
        }
        if (sg.zb > 0) {                    // early part of a cut chain: hand the adjoint wavefield to CTA c - 1
            float4* dst = p.handoff + (size_t)c * p.slot_f4;
            for (int i = t; i < (int)(C::PSI / 16); i += C::NTHREADS) dst[i] = ((const float4*)psi)[i];
#ifndef ADM_HOST_EMU
            __threadfence();
#endif
            __syncthreads();
            if (t == 0) flag_set(p.flags + c);
        }
    }
    if (t == 0) bulk_wait_all();      // shared memory must stay valid until the last reduce has read it
}
