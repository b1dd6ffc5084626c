// Fused multislice forward / adjoint kernels (sm_100a).
//
// One CTA owns one wavefield (pattern b, probe mode m).  The wavefield psi[P][P] lives in shared memory for the
// whole traversal of the Nz slices.  The object enters as the TRANSMISSION VOLUME T' = T - 1 (slice-major
// [Nz][Y][Xp][2], written once per step by adm_transmission_fwd, so no exp/sincos runs in the per-slice loop and the
// same kernel serves delta_beta and real_imag objects).  Per slice the only HBM traffic is
//   forward : one TMA tile load of T' (cp.async.bulk.tensor.3d, out-of-object voxels zero-filled = vacuum),
//             prefetched one slice ahead and signalled on an mbarrier, + the psi checkpoint store;
//   adjoint : one cp.async.bulk of the psi checkpoint (prefetched), the T' tile (L2-resident, plain loads issued
//             ahead of the butterfly) and ONE cp.reduce.async.bulk.tensor.3d that scatter-ADDS the [P][P] tile of
//             dL/dT into the gradient volume (reduction done by the L2, overhang clipped by the TMA unit).
//
// 2-D FFT of size P = 8*R2 (R2 = 8 or 9) as an in-place mixed-radix Cooley-Tukey over shared memory:
//   1-D forward:  pass A  radix-8 over n1 (x[R2 n1 + n2], stride R2) + twiddle W_P^{n2 k1}, stored at R2 k1 + n2
//                 pass B  radix-R2 over n2 (contiguous R2 k1 + n2)               -> X[k1 + 8 k2] at R2 k1 + k2
//   1-D inverse:  pass B^-1 then pass A^-1 on the same index sets, conjugate twiddles, unnormalised.
//   The spectrum therefore sits in a permuted order, perm(p) = p / R2 + 8 (p % R2); the transfer function is
//   pre-permuted (adm_prepare_transfer_function), so no reordering pass exists.
// Because pass B, the H multiply and pass B^-1 act on the same R2 points, and pass A^-1, the slice modulation
// and the next slice's pass A act on the same 8 points, a slice costs SIX shared-memory round trips:
//   P1 [Ay^-1 . (1 + T') . Ay]   P2 By   P3 Ax   P4 [Bx . H . Bx^-1]   P5 Ax^-1   P6 By^-1
// y-passes run with lanes along x (row stride S = P+1 complex keeps x-passes, lanes along y, conflict free).
//
// Reference behaviour reproduced: adorym/propagate.py:208-258 (slice loop, last slice not propagated :253),
// :241/:246 transmission (precomputed), :255 convolution, :263-280 detector transform; forward_model.py:313-331 +
// util.py:1327-1351 gather with vacuum padding.  The adjoint replaces the autograd tape (wrappers.py:300-331).
#include <stdlib.h>
#include <stdio.h>
#include "fft_passes.cuh"
#include "tma.cuh"
#include "../../include/adorym_b200.h"
#include "ms_params.cuh"


// shared-memory carve-up (tile first: TMA destinations / sources need 128-byte alignment)
template <class C> struct MsSmem {
    static constexpr int TP = C::P + 2;      // tile row pitch: P columns + 2 (even-column TMA start, see tma.cuh)
    static constexpr size_t TILE = sizeof(cpx) * C::P * TP;
    static constexpr size_t PSI = sizeof(cpx) * C::P * C::S;
    static constexpr size_t TW = sizeof(cpx) * C::P;
    static constexpr size_t HS = sizeof(cpx) * 2 * HsLayout<C>::N;     // separable transfer function tables hy | hx (per thread row)
    static constexpr size_t BYTES = TILE + PSI + TW + HS + 16;
    static constexpr size_t BYTES_NOTILE = PSI + TW + HS + 16;     // forward variant that reads T' with plain loads
    static constexpr int PSI_F4 = (int)(PSI / 16);           // the psi buffer as float4 (hand-off copies)
    static_assert(PSI % 16 == 0 && TILE % 16 == 0, "16-byte granularity");
};

// ---- persistent balanced schedule ---------------------------------------------------------------------------------
// A launch owns nwf = B * M independent wavefields, each a strictly sequential chain of Nz slice steps.  With one
// CTA per wavefield and two CTAs resident per SM, 529 wavefields on 296 slots run as 1.79 "waves": the second wave
// leaves a fifth of the slots idle.  Instead the nwf * Nz slice steps are laid end to end and cut into gridDim.x
// intervals of T steps (McNaughton's wrap-around rule): every CTA does the same number of slice steps, and a
// wavefield cut by an interval boundary is handed from one CTA to the other through a raw copy of the shared-memory
// wavefield (42 KB, once per CTA) and a flag.  The CTA that holds the EARLY part of a cut chain processes it FIRST, the
// CTA that holds the late part processes it LAST (T >= Nz), so the consumer practically never waits; the dependency
// always points to a lower blockIdx, and the grid never exceeds the number of co-resident CTAs.
struct MsSeg { int wf, zb, ze; };
// c: logical CTA index, i: position in this CTA's processing order (0 = first).  fwd: chains run z upwards, the cut
// wavefield at the END of the interval is its early part -> processed first; bwd: chains run z downwards, the cut
// wavefield at the START of the interval is its early part -> processed first.
ADM_DEV int ms_nseg(const MsParams& p, int c, int& wf_first, int& wf_last, long long& s0, long long& s1) {
    const long long total = (long long)p.B * p.M * p.Nz;
    if (!p.split) { s0 = (long long)c * p.Nz; s1 = s0 + p.Nz; }
    else { s0 = (long long)c * p.T; s1 = s0 + p.T; if (s1 > total) s1 = total; }
    wf_first = (int)(s0 / p.Nz); wf_last = (int)((s1 - 1) / p.Nz);
    return wf_last - wf_first + 1;
}
ADM_DEV MsSeg ms_seg(const MsParams& p, int wf, long long s0, long long s1) {
    MsSeg s; s.wf = wf;
    long long lo = s0 - (long long)wf * p.Nz, hi = s1 - (long long)wf * p.Nz;
    s.zb = lo > 0 ? (int)lo : 0; s.ze = hi < p.Nz ? (int)hi : p.Nz;
    return s;
}
ADM_DEV void flag_set(int* f) {
#ifndef ADM_HOST_EMU
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(f), "r"(1) : "memory");
#else
    std::atomic_ref<int>(*f).store(1);
#endif
}
ADM_DEV void flag_wait_and_clear(int* f) {
#ifndef ADM_HOST_EMU
    int v;
    do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(f) : "memory"); } while (v == 0);
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;\n" ::"l"(f), "r"(0) : "memory");
#else
    std::atomic_ref<int> r(*f);
    while (r.load() == 0) std::this_thread::yield();
    r.store(0);
#endif
}
// raw copy of the shared-memory wavefield to / from a hand-off slot (all threads; barriers by the caller)
template <class C> ADM_DEV void handoff_store(const MsParams& p, int slot, const void* psi) {
    float4* dst = p.handoff + (size_t)slot * p.slot_f4;
    const float4* src = (const float4*)psi;
    for (int i = threadIdx.x; i < MsSmem<C>::PSI_F4; i += C::NTHREADS) dst[i] = src[i];
#ifndef ADM_HOST_EMU
    __threadfence();
#endif
}
template <class C> ADM_DEV void handoff_load(const MsParams& p, int slot, void* psi) {
    const float4* src = p.handoff + (size_t)slot * p.slot_f4;
    float4* dst = (float4*)psi;
    for (int i = threadIdx.x; i < MsSmem<C>::PSI_F4; i += C::NTHREADS) {
#ifndef ADM_HOST_EMU
        dst[i] = __ldcg(src + i);
#else
        dst[i] = src[i];
#endif
    }
}

#include "ms3_kernels.cuh"

// ================================================================================================
// Forward
// ================================================================================================
// TILE = true : T' tile by TMA into shared memory, prefetched one slice ahead (86 KB of shared memory, 2 CTAs per SM).
// TILE = false: T' by plain coalesced loads at the top of P1 (the volume is L2 resident), 44 KB of shared memory: with
//               <= 96 registers THREE CTAs are resident per SM (21 warps instead of 14 to hide barrier / dependency stalls).
template <class C, int MINB, bool TILE, class T> __global__ void __launch_bounds__(C::NTHREADS, MINB) k_multislice_fwd(const ADM_GRID_CONSTANT AdmTileMap tmap, MsParams p) {
    ADM_SMEM_DECL
    constexpr int P = C::P, R2 = C::R2, S = C::S;
    constexpr unsigned TILE_BYTES = (unsigned)MsSmem<C>::TILE;
    constexpr int TP = MsSmem<C>::TP;
    T* tile = (T*)smem_raw;
    T* psi = TILE ? tile + P * TP : (T*)smem_raw;
    T* tw = psi + P * S;
    T* hs = tw + P;
    adm_mbar_t* bar = (adm_mbar_t*)(hs + 2 * HsLayout<C>::N);
    const int t = threadIdx.x;
    const bool active = t < C::NACT;
    const int q = t / P, l = t % P;
    const Idx<C> I(q);
    const int c = blockIdx.x;               // logical CTA index (dependencies point to c - 1)

    init_twiddles<C>(tw, t);
    if (p.h_separable) hs_fill<C>(hs, (const T*)p.Hp, t);
    if (TILE && t == 0) mbar_init(bar, 1);
    unsigned ph = 0;                        // parity of the next tile completion

    int wf_first, wf_last;
    long long s0, s1;
    ms_nseg(p, c, wf_first, wf_last, s0, s1);
    for (int wfi = wf_last; wfi >= wf_first; --wfi) {
        const MsSeg sg = ms_seg(p, wfi, s0, s1);
        const int wf = sg.wf;               // wavefield index b * M + m
        const int b = wf / p.M, m = wf % p.M;
        const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
        const int pxa = px & ~1, lt = l + (px & 1);     // even TMA start column, this thread's column inside the tile
        const T* probe = (const T*)p.probe + (size_t)(p.probe_per_pattern ? wf : m) * P * P;
        T* ckpt = (T*)p.ckpt + (size_t)wf * p.Nz * P * TP;
        T* det = (T*)p.det + (size_t)wf * P * P;

        __syncthreads();                    // the previous segment is done with the tile and the wavefield
        if (TILE && t == 0) {               // prefetch the T' tile of the first slice of the segment
            mbar_expect_tx(bar, TILE_BYTES);
            tma_load_tile(tile, &tmap, 2 * pxa, py, sg.zb, bar);
        }
        // TILE = false: this thread's column of the transmission volume
        const int gx = px + l;
        const bool x_in = (gx >= 0) && (gx < p.X);
        const T* tcol = (const T*)p.tvol + (x_in ? gx : 0);
        const size_t slice_stride = (size_t)p.Y * p.Xp;
        if (sg.zb > 0) {                    // late part of a cut chain: the wavefield comes from CTA c - 1
            if (t == 0) flag_wait_and_clear(p.flags + c);
            __syncthreads();
            handoff_load<C>(p, c, psi);
            __syncthreads();
        }

        for (int z = sg.zb; z < sg.ze; ++z) {
            const bool last = (z == p.Nz - 1);
            const bool go_on = !last || p.free_prop != ADM_PROP_NONE;
            // per-gap transfer functions (sparse multislice): tables of the gap z -> z+1; the previous slice's last reader
            // (its x pass) is two barriers back, the first reader (P2') comes after the barrier that closes P1
            if (!last && p.h_per_slice) hs_fill<C>(hs, (const T*)p.Hp + (size_t)z * 2 * P, t);
            // ---- P1: [Ay^-1] . checkpoint . modulation . [Ay] ------------------------------------------------
            const T* tz = tcol + (size_t)z * slice_stride;
            if (active) {
                if (TILE) mbar_wait(bar, ph);
                ADM_FOR_GROUPS_A(gi) {
                    const int g = I.group(gi);
                    T v[8], tp[8];
                    if (!TILE) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int gy = py + I.a(j, gi);
                            const bool inb = x_in && (unsigned)gy < (unsigned)p.Y;
                            tp[j] = inb ? tz[(unsigned)(gy * p.Xp)] : tmk<T>(0.f, 0.f);
                        }
                    }
                    if (z == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = probe[I.a(j, gi) * P + l];
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = psi[I.a(j, gi) * S + l];
                        dft8<true>(v);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int y = I.a(j, gi);
                        ckpt[(size_t)z * P * TP + y * TP + lt] = v[j];
                        v[j] = modulate(v[j], TILE ? tile[y * TP + lt] : tp[j]);
                    }
                    if (go_on) {
                        if (last && p.ff_conj) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = cconj(v[j]);
                        }
                        dft8<false>(v);
                        if constexpr (!C::PFA) {
#pragma unroll
                            for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw[g * k]);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) psi[I.a(k, gi) * S + l] = v[k];
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) det[I.a(j, gi) * P + l] = v[j];   // exit wave, natural order
                    }
                }
            }
            ph ^= 1;
            __syncthreads();
            if (TILE && t == 0 && z + 1 < sg.ze) {   // every thread has consumed tile(z): prefetch tile(z+1) behind P2..P5
                mbar_expect_tx(bar, TILE_BYTES);
                tma_load_tile(tile, &tmap, 2 * pxa, py, z + 1, bar);
            }
            if (!go_on) break;
            if (last && p.free_prop == ADM_PROP_FAR) {
                // far field: finish the forward 2-D FFT (By, Ax, Bx) and store fftshift-ed, natural order
                if (active) pass_B_fwd<C, false>(psi, q, l);
                __syncthreads();
                if (active) pass_A_fwd<C, true>(psi, tw, q, l);
                __syncthreads();
                if (active) {
                    const int fy = (perm_freq<C>(l) + P / 2) % P;
                    ADM_FOR_GROUPS_B(gi) {
                        T v[R2];
#pragma unroll
                        for (int n = 0; n < R2; ++n) v[n] = psi[l * S + I.b(n, gi)];
                        dftr<R2, false>(v);
#pragma unroll
                        for (int k = 0; k < R2; ++k) {
                            int fx = (perm_freq<C>(I.b(k, gi)) + P / 2) % P;
                            T o = cscale(v[k], p.ff_scale);
                            det[fy * P + fx] = p.ff_conj ? cconj(o) : o;
                        }
                    }
                }
                break;
            }
            if (!last && p.h_separable) propagate_core_sep5<C>(psi, tw, hs, q, l, active);      // five passes per slice
            else propagate_core<C>(psi, tw, (const T*)(last ? p.Hf : p.Hp), q, l, active);
            if (last) {
                // near field: final Ay^-1, natural order store
                if (active) {
                    ADM_FOR_GROUPS_A(gi) {
                        T v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = psi[I.a(j, gi) * S + l];
                        dft8<true>(v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) det[I.a(j, gi) * P + l] = v[j];
                    }
                }
            }
        }
        if (sg.ze < p.Nz) {                 // early part of a cut chain: hand the wavefield to CTA c + 1
            handoff_store<C>(p, c + 1, psi);
            __syncthreads();
            if (t == 0) flag_set(p.flags + c + 1);
        }
    }
}

// ================================================================================================
// Adjoint
// ================================================================================================
template <class C, int MINB, class T, bool GCK = false> __global__ void __launch_bounds__(C::NTHREADS, MINB) k_multislice_bwd(const ADM_GRID_CONSTANT AdmTileMap gmap, MsParams p) {
    ADM_SMEM_DECL
    constexpr int P = C::P, R2 = C::R2, S = C::S;
    constexpr unsigned TILE_BYTES = (unsigned)MsSmem<C>::TILE;
    constexpr int TP = MsSmem<C>::TP;
    T* ckb = (T*)smem_raw;          // psi checkpoint of the slice (bulk copy); overwritten in place by dL/dT
    T* psi = ckb + P * TP;            // running adjoint wavefield g
    T* tw = psi + P * S;
    T* hs = tw + P;
    adm_mbar_t* bar = (adm_mbar_t*)(hs + 2 * HsLayout<C>::N);
    const int t = threadIdx.x;
    const bool active = t < C::NACT;
    const int q = t / P, l = t % P;
    const Idx<C> I(q);
    const int c = gridDim.x - 1 - blockIdx.x;   // logical CTA index; the chains run downwards, so CTA c depends on the
                                                // FIRST segment of CTA c + 1 = a lower blockIdx

    init_twiddles<C>(tw, t);
    if (p.h_separable) hs_fill<C>(hs, (const T*)p.Hp, t);
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
        const int pxa = px & ~1, lt = l + (px & 1);
        const int gx = px + l;
        const bool x_in = (gx >= 0) && (gx < p.X);
        const T* ckpt = (T*)p.ckpt + (size_t)wf * p.Nz * P * TP;
        const T* gdet = (const T*)p.det + (size_t)wf * P * P;
        const size_t slice_stride = (size_t)p.Y * p.Xp;
        const T* tcol = (const T*)p.tvol + (x_in ? gx : 0);     // this thread's column of the transmission volume
        // T' of one slice for this thread's P1' groups (L2-resident volume, plain loads with 32-bit offsets from the
        // per-slice column pointer).  The loads are ISSUED one barrier early -- behind the last pass of the previous
        // slice -- so that their L2 latency (~700 cycles, the adjoint's former top stall) overlaps the barrier wait.
        T tpa[C::NGA][8];
        auto load_tp = [&](int zz) {
            const T* tzz = tcol + (size_t)zz * slice_stride;
            ADM_FOR_GROUPS_A(gi) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int gy = py + I.a(j, gi);
                    const bool inb = x_in && (unsigned)gy < (unsigned)p.Y;
                    tpa[gi][j] = inb ? tzz[(unsigned)(gy * p.Xp)] : tmk<T>(0.f, 0.f);
                }
            }
        };
        if (active) load_tp(sg.ze - 1);

        __syncthreads();                    // the previous segment is done with the wavefield
        if (t == 0) {   // prefetch the checkpoint of the first (highest) slice of the segment
            bulk_wait_read_all();           // ... and its last reduce has drained the tile
            mbar_expect_tx(bar, TILE_BYTES);
            bulk_load(ckb, ckpt + (size_t)(sg.ze - 1) * P * TP, TILE_BYTES, bar);
        }

        if (sg.ze < p.Nz) {
            // late part of a cut chain: the adjoint wavefield comes from CTA c + 1
            if (t == 0) flag_wait_and_clear(p.flags + c + 1);
            __syncthreads();
            handoff_load<C>(p, c + 1, psi);
            __syncthreads();
        } else if (p.free_prop == ADM_PROP_FAR) {
            // ---- adjoint of the detector transform -------------------------------------------------------
            if (active) {
                const int fy = (perm_freq<C>(l) + P / 2) % P;
                ADM_FOR_GROUPS_B(gi) {
                    const int g = I.group(gi);
                    T v[R2];
#pragma unroll
                    for (int k = 0; k < R2; ++k) {
                        int fx = (perm_freq<C>(I.b(k, gi)) + P / 2) % P;
                        T o = cscale(gdet[fy * P + fx], p.ff_scale);
                        v[k] = p.ff_conj ? cconj(o) : o;
                    }
                    dftr<R2, true>(v);
                    if constexpr (!C::PFA) {
#pragma unroll
                        for (int n = 1; n < R2; ++n) v[n] = cmulc(v[n], tw[g * n]);
                    }
#pragma unroll
                    for (int n = 0; n < R2; ++n) psi[l * S + I.b(n, gi)] = v[n];
                }
            }
            __syncthreads();
            if (active) pass_A_inv<C, true>(psi, q, l);
            __syncthreads();
            if (active) pass_B_inv<C, false>(psi, tw, q, l);
            __syncthreads();
        } else if (p.free_prop == ADM_PROP_NEAR) {
            if (active) {
                ADM_FOR_GROUPS_A(gi) {
                    const int g = I.group(gi);
                    T v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = gdet[I.a(j, gi) * P + l];
                    dft8<false>(v);
                    if constexpr (!C::PFA) {
#pragma unroll
                        for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw[g * k]);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) psi[I.a(k, gi) * S + l] = v[k];
                }
            }
            __syncthreads();
            propagate_core<C>(psi, tw, (const T*)p.Hf, q, l, active);
        }

        for (int z = sg.ze - 1; z >= sg.zb; --z, phase ^= 1) {
            const bool last = (z == p.Nz - 1);
            // per-gap transfer functions: tables of the gap z-1 -> z (read after the barrier that closes P1')
            if (z > 0 && p.h_per_slice) hs_fill<C>(hs, (const T*)p.Hp + (size_t)(z - 1) * 2 * P, t);
            // ---- P1': [Ay^-1] . dL/dT = g' conj(psi_z) . g = g' conj(1 + T') . [Ay] --------------------------
            if (active) {
                ADM_FOR_GROUPS_A(gi) {
                    const int g = I.group(gi);
                    T v[8];
                    if (last && p.free_prop == ADM_PROP_NONE) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = gdet[I.a(j, gi) * P + l];
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = psi[I.a(j, gi) * S + l];
                        dft8<true>(v);
                        if (last && p.ff_conj) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = cconj(v[j]);
                        }
                    }
                    if (gi == 0) {
                        mbar_wait(bar, phase);
                        // the two pad columns of the tile are never written by the forward pass: add zeros there
                        if (t < 2 * P) ckb[(t >> 1) * TP + ((t & 1) ? P + 1 : ((px & 1) ? 0 : P))] = tmk<T>(0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int y = I.a(j, gi);
                        const T w = ckb[y * TP + lt];               // psi_z (wave entering the slice)
                        ckb[y * TP + lt] = cmulc(v[j], w);            // dL/dT = g' conj(psi_z)
                        v[j] = modulate_conj(v[j], tpa[gi][j]);       // g_psi = g' conj(T)
                    }
                    if constexpr (GCK) {           // refinable slice positions: dL/d psi_z for the host-side d/d gap
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            ((T*)p.g_ckpt)[((size_t)wf * p.Nz + z) * P * P + I.a(j, gi) * P + l] = v[j];
                    }
                    if (z > 0) {
                        dft8<false>(v);
                        if constexpr (!C::PFA) {
#pragma unroll
                            for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw[g * k]);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) psi[I.a(k, gi) * S + l] = v[k];
                    } else if (p.g_probe != nullptr) {
                        if (p.probe_per_pattern) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) ((T*)p.g_probe)[(size_t)wf * P * P + I.a(j, gi) * P + l] = v[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                red_add2(&p.g_probe[(size_t)m * P * P + I.a(j, gi) * P + l], cx(v[j]), cy(v[j]));
                        }
                    }
                }
                fence_proxy_async();      // the dL/dT tile is about to be read by the TMA unit
            }
            __syncthreads();
            if (t == 0 && pxa + P >= 0 && py + P >= 0) {   // (a patch lying entirely outside contributes nothing)
                tma_reduce_add_tile(&gmap, 2 * (pxa + P), py + P, z, ckb);   // margined volume: coordinates never negative
                bulk_commit();
            }
            if (z == 0) { phase ^= 1; break; }
            // adjoint propagation to slice z-1.  Separable H: five passes [By conj(hy) By^-1] [Ax] [Bx conj(hx) Bx^-1] [Ax^-1]
            // (+ the Ay^-1 at the head of the next P1'); general H: six.
            if (active) {
                if (p.h_separable) pass_BHB_y_sep<C>(psi, tw, hs, q, l);
                else pass_B_fwd<C, false>(psi, q, l);
            }
            __syncthreads();
            if (t == 0 && z > sg.zb) {   // the reduce has drained the tile: prefetch the checkpoint of slice z-1 behind the x passes
                bulk_wait_read_all();
                mbar_expect_tx(bar, TILE_BYTES);
                bulk_load(ckb, ckpt + (size_t)(z - 1) * P * TP, TILE_BYTES, bar);
            }
            if (active) pass_A_fwd<C, true>(psi, tw, q, l);
            __syncthreads();
            if (active) {
                if (p.h_separable) pass_BHB_x_sep1<C>(psi, tw, hs + HsLayout<C>::N, q, l);
                else pass_BHB_x<C>(psi, tw, (const T*)p.Hp, q, l);
            }
            __syncthreads();
            if (active) pass_A_inv<C, true>(psi, q, l);
            if (p.h_separable && active && z > sg.zb) load_tp(z - 1);      // T' of the next slice, one barrier early
            __syncthreads();
            if (!p.h_separable) {
                if (active) pass_B_inv<C, false>(psi, tw, q, l);
                if (active && z > sg.zb) load_tp(z - 1);
                __syncthreads();
            }
        }
        if (sg.zb > 0) {                    // early part of a cut chain: hand the adjoint wavefield to CTA c - 1
            handoff_store<C>(p, c, psi);
            __syncthreads();
            if (t == 0) flag_set(p.flags + c);
        }
    }
    if (t == 0) bulk_wait_all();      // shared memory must stay valid until the last reduce has read it
}

// ================================================================================================
// transfer-function preparation
// ================================================================================================
template <class C> __global__ void k_prepare_H(const cpx* H, int conj, float scale, cpx* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C::P * C::P) return;
    int pxp = i / C::P, pyp = i % C::P;     // out[px][py]
    cpx h = H[perm_freq<C>(pyp) * C::P + perm_freq<C>(pxp)];
    if (conj) h = cconj(h);
    out[i] = cscale(h, scale);
}

// ================================================================================================
// host side
// ================================================================================================
// kernel-shape variant selection for tuning runs (0 = default)
static int adm_variant(const char* name) {
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
}

// three-pass kernels (ms3_kernels.cuh) for a separable transfer function (P = 72 and 64); ADM_MS3=0 selects the five-pass kernels
static bool ms3_wanted(const adm_multislice_cfg* cfg, const MsParams& p) {
    const char* e = getenv("ADM_MS3");
    if (e && atoi(e) == 0) return false;
    return (cfg->P == 72 || cfg->P == 64) && p.h_separable && p.Nz > 1 && (p.free_prop != ADM_PROP_NEAR || p.hfree_separable);
}

// number of CTAs of `kern` that are co-resident on the device (SM count x occupancy); ADM_MS_GRID overrides it (tests)
template <class K> static int ms_resident_ctas(K kern, int threads, size_t smem) {
    int forced = adm_variant("ADM_MS_GRID");
    if (forced > 0) return forced;
#ifndef ADM_HOST_EMU
    // cache per kernel (K is the function-pointer TYPE, shared by all kernels of this signature: key on the pointer)
    static const void* keys[16];
    static int vals[16];
    static int nkeys = 0;
    for (int i = 0; i < nkeys; ++i)
        if (keys[i] == (const void*)kern) return vals[i];
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess) return 0;
    const int res = sms * occ;
    if (nkeys < 16) { keys[nkeys] = (const void*)kern; vals[nkeys] = res; ++nkeys; }
    if (getenv("ADM_MS_VERBOSE"))
        fprintf(stderr, "[adm] multislice kernel: %d threads, %zu B smem -> %d CTAs per SM, %d resident\n", threads, smem, occ, res);
    return res;
#else
    return 8;
#endif
}

#define ADM_MS_MAX_SLOTS 1024
// workspace: [ADM_MS_MAX_SLOTS] int flags | pad to 256 B | [ADM_MS_MAX_SLOTS][slot] raw wavefield copies
static size_t ms_slot_bytes(int P) { return (size_t)P * (P + 1) * sizeof(cpx); }

// schedule: one wavefield per CTA while they all fit; otherwise (and only with a workspace) equal intervals of slice steps
static int ms_schedule(MsParams& p, int cap, void* workspace, int P, int& grid) {
    const int nwf = p.B * p.M;
    p.split = 0; p.T = p.Nz; p.flags = nullptr; p.handoff = nullptr; p.slot_f4 = (int)(ms_slot_bytes(P) / 16);
    grid = nwf;
    if (cap <= 0) return ADM_ERR_CUDA;
    if (workspace != nullptr && nwf > cap && cap + 1 <= ADM_MS_MAX_SLOTS) {
        const long long total = (long long)nwf * p.Nz;
        const long long T = (total + cap - 1) / cap;          // >= Nz because nwf > cap
        p.split = 1; p.T = (int)T;
        grid = (int)((total + T - 1) / T);
        p.flags = (int*)workspace;
        p.handoff = (float4*)((char*)workspace + 4096);
    }
    return ADM_OK;
}

extern "C" long long adm_multislice_workspace_bytes(int P) {
    if (P != 64 && P != 72) return 0;
    return 4096 + (long long)ADM_MS_MAX_SLOTS * (long long)ms_slot_bytes(P);
}

static int fill_params(const adm_multislice_cfg* c, MsParams& p) {
    if (!c) return ADM_ERR_ARG;
    if (c->P != 64 && c->P != 72) return ADM_ERR_UNSUPPORTED;
    if (c->n_slices < 1 || c->B < 1 || c->M < 1 || c->Y < 1 || c->X < 1) return ADM_ERR_ARG;
    if (c->sign_convention != 1 && c->sign_convention != -1) return ADM_ERR_ARG;
    if (c->unknown_type != ADM_DELTA_BETA && c->unknown_type != ADM_REAL_IMAG) return ADM_ERR_ARG;
    if (c->free_prop < ADM_PROP_NONE || c->free_prop > ADM_PROP_NEAR) return ADM_ERR_ARG;
    p.Nz = c->n_slices; p.Y = c->Y; p.X = c->X; p.Xp = ADM_TVOL_PITCH(c->X); p.B = c->B; p.M = c->M;
    p.free_prop = c->free_prop; p.probe_per_pattern = c->probe_per_pattern; p.h_separable = c->h_separable ? 1 : 0;
    p.h_per_slice = c->h_per_slice ? 1 : 0;
    p.hfree_separable = c->hfree_separable ? 1 : 0;
    if (p.h_per_slice && !p.h_separable) return ADM_ERR_UNSUPPORTED;
    // far field: sign +1 -> fft2 (unnormalised or 1/P), sign -1 -> ifft2 (1/P^2 or 1/P)   (propagate.py:267-270)
    float Pf = (float)c->P;
    p.ff_conj = (c->sign_convention == -1 && c->free_prop == ADM_PROP_FAR) ? 1 : 0;
    if (c->normalize_fft) p.ff_scale = 1.0f / Pf;
    else p.ff_scale = (c->sign_convention == -1) ? 1.0f / (Pf * Pf) : 1.0f;
    return ADM_OK;
}

extern "C" int adm_abi_version(void) { return ADM_ABI_VERSION; }

extern "C" int adm_prepare_transfer_function(int P, const float* H, int conj, float scale, float* out, void* stream) {
    if (!H || !out) return ADM_ERR_ARG;
    if (P == 72) ADM_LAUNCH(k_prepare_H<Cfg72>, (72 * 72 + 255) / 256, 256, 0, stream, (const cpx*)H, conj, scale, (cpx*)out);
    else if (P == 64) ADM_LAUNCH(k_prepare_H<Cfg64>, (64 * 64 + 255) / 256, 256, 0, stream, (const cpx*)H, conj, scale, (cpx*)out);
    else return ADM_ERR_UNSUPPORTED;
    return ADM_OK;
}

// separable variant: out = [hy_perm (P) | scale * hx_perm (P)], complex; conj optional
template <class C> __global__ void k_prepare_H_sep(const cpx* hy, const cpx* hx, int conj, float scale, cpx* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * C::P) return;
    cpx h = (i < C::P) ? hy[perm_freq<C>(i)] : cscale(hx[perm_freq<C>(i - C::P)], scale);
    out[i] = conj ? cconj(h) : h;
}

extern "C" int adm_prepare_transfer_function_separable(int P, const float* hy, const float* hx, int conj, float scale,
                                                       float* out, void* stream) {
    if (!hy || !hx || !out) return ADM_ERR_ARG;
    if (P == 72) ADM_LAUNCH(k_prepare_H_sep<Cfg72>, 1, 256, 0, stream, (const cpx*)hy, (const cpx*)hx, conj, scale, (cpx*)out);
    else if (P == 64) ADM_LAUNCH(k_prepare_H_sep<Cfg64>, 1, 256, 0, stream, (const cpx*)hy, (const cpx*)hx, conj, scale, (cpx*)out);
    else return ADM_ERR_UNSUPPORTED;
    return ADM_OK;
}

extern "C" int adm_multislice_fwd(const adm_multislice_cfg* cfg, const float* tvol, const float* probe, const int* pos,
                                  const float* Hprep, const float* Hfree, float* psi_ckpt, float* det, void* workspace,
                                  void* stream) {
    MsParams p;
    int rc = fill_params(cfg, p);
    if (rc) return rc;
    if (!tvol || !probe || !pos || !psi_ckpt || !det) return ADM_ERR_ARG;
    if (cfg->n_slices > 1 && !Hprep) return ADM_ERR_ARG;
    if (cfg->free_prop == ADM_PROP_NEAR && !Hfree) return ADM_ERR_ARG;
    p.tvol = (const cpx*)tvol; p.probe = (const cpx*)probe; p.pos = pos;
    p.Hp = (const cpx*)Hprep; p.Hf = (const cpx*)Hfree;
    p.ckpt = (cpx*)psi_ckpt; p.det = (cpx*)det; p.g_probe = nullptr; p.g_ckpt = nullptr;
    AdmTileMap map;
    if (adm_make_tile_map(&map, tvol, p.Nz, p.Y, p.X, p.Xp, cfg->P, cfg->P + 2)) return ADM_ERR_CUDA;
    int var = adm_variant("ADM_FWD_VARIANT");
    if (var == 0 && ms3_wanted(cfg, p)) {
#define ADM_MS3_LAUNCH(KERN, CFG)                                                                                      \
    do {                                                                                                            \
        if (set_smem(KERN<cpk, CFG>, CFG::BYTES)) return ADM_ERR_CUDA;                                              \
        int grid_;                                                                                                  \
        if (ms_schedule(p, ms_resident_ctas(KERN<cpk, CFG>, CFG::NTHREADS, CFG::BYTES), workspace, CFG::P, grid_))  \
            return ADM_ERR_CUDA;                                                                                    \
        ADM_LAUNCH((KERN<cpk, CFG>), grid_, CFG::NTHREADS, CFG::BYTES, stream, map, p);                             \
    } while (0)
        // 9 warps (M pass in two balanced rounds) when the SMs are shared by many wavefields; 7 warps for a handful of
        // wavefields, where the latency of one slice step counts (reference demo schedule, minibatch 2: 2.76 vs 2.88 ms per step)
        if (cfg->P == 72 && (p.B * p.M > 148 || adm_variant("ADM_MS3_FWD") == 2)) ADM_MS3_LAUNCH(k_ms3_fwd, Ms3w);      // (2: tests)
        else if (cfg->P == 72) ADM_MS3_LAUNCH(k_ms3_fwd, Ms3);
        else ADM_MS3_LAUNCH(k_ms3_fwd, Ms3p64);
        return ADM_OK;
    }
#define ADM_FWD(CFG, MINB, TILE, CT)                                                                                  \
    do {                                                                                                           \
        const size_t sm_ = TILE ? MsSmem<CFG>::BYTES : MsSmem<CFG>::BYTES_NOTILE;                                  \
        if (set_smem(k_multislice_fwd<CFG, MINB, TILE, CT>, sm_)) return ADM_ERR_CUDA;                             \
        int grid_;                                                                                                 \
        if (ms_schedule(p, ms_resident_ctas(k_multislice_fwd<CFG, MINB, TILE, CT>, CFG::NTHREADS, sm_), workspace, \
                        CFG::P, grid_))                                                                            \
            return ADM_ERR_CUDA;                                                                                   \
        ADM_LAUNCH((k_multislice_fwd<CFG, MINB, TILE, CT>), grid_, CFG::NTHREADS, sm_, stream, map, p);            \
    } while (0)
    if (cfg->P == 72) {
        if (var == 4) ADM_FWD(Cfg72, 2, true, cpx);      // scalar butterflies (one FADD / FFMA per component)
        else if (var == 3) ADM_FWD(Cfg72, 3, false, cpk);   // no T' tile, three CTAs per SM
        else if (var == 2) ADM_FWD(Cfg72, 2, false, cpk);   // no T' tile, two CTAs per SM
        else ADM_FWD(Cfg72, 2, true, cpk);               // packed complex arithmetic (cpk.cuh)
    } else {
        if (var == 4) ADM_FWD(Cfg64, 2, true, cpx);
        else ADM_FWD(Cfg64, 2, true, cpk);
    }
    return ADM_OK;
}

static int ms_bwd_impl(const adm_multislice_cfg* cfg, const float* tvol, const int* pos, const float* HprepC,
                       const float* HfreeC, const float* psi_ckpt, const float* g_det, float* g_tvol, float* g_probe,
                       float* g_psi_ckpt, void* workspace, void* stream);

extern "C" int adm_multislice_bwd(const adm_multislice_cfg* cfg, const float* tvol, const int* pos, const float* HprepC,
                                  const float* HfreeC, const float* psi_ckpt, const float* g_det, float* g_tvol,
                                  float* g_probe, void* workspace, void* stream) {
    return ms_bwd_impl(cfg, tvol, pos, HprepC, HfreeC, psi_ckpt, g_det, g_tvol, g_probe, nullptr, workspace, stream);
}

extern "C" int adm_multislice_bwd_gpsi(const adm_multislice_cfg* cfg, const float* tvol, const int* pos, const float* HprepC,
                                       const float* HfreeC, const float* psi_ckpt, const float* g_det, float* g_tvol,
                                       float* g_probe, float* g_psi_ckpt, void* workspace, void* stream) {
    if (!g_psi_ckpt) return ADM_ERR_ARG;
    return ms_bwd_impl(cfg, tvol, pos, HprepC, HfreeC, psi_ckpt, g_det, g_tvol, g_probe, g_psi_ckpt, workspace, stream);
}

static int ms_bwd_impl(const adm_multislice_cfg* cfg, const float* tvol, const int* pos, const float* HprepC,
                       const float* HfreeC, const float* psi_ckpt, const float* g_det, float* g_tvol, float* g_probe,
                       float* g_psi_ckpt, void* workspace, void* stream) {
    MsParams p;
    int rc = fill_params(cfg, p);
    if (rc) return rc;
    if (!tvol || !pos || !psi_ckpt || !g_det || !g_tvol) return ADM_ERR_ARG;
    if (cfg->n_slices > 1 && !HprepC) return ADM_ERR_ARG;
    if (cfg->free_prop == ADM_PROP_NEAR && !HfreeC) return ADM_ERR_ARG;
    p.tvol = (const cpx*)tvol; p.probe = nullptr; p.pos = pos;
    p.Hp = (const cpx*)HprepC; p.Hf = (const cpx*)HfreeC;
    p.ckpt = (cpx*)psi_ckpt; p.det = (cpx*)g_det; p.g_probe = (cpx*)g_probe; p.g_ckpt = (cpx*)g_psi_ckpt;
    AdmTileMap map;
    if (adm_make_tile_map(&map, g_tvol, p.Nz, ADM_GVOL_ROWS(p.Y, cfg->P), p.X + ADM_GVOL_MARGIN(cfg->P),
                          ADM_GVOL_PITCH(p.X, cfg->P), cfg->P, cfg->P + 2))
        return ADM_ERR_CUDA;
    int var = adm_variant("ADM_BWD_VARIANT");
    if (var == 0 && p.g_ckpt == nullptr && ms3_wanted(cfg, p)) {
        if (cfg->P == 72) ADM_MS3_LAUNCH(k_ms3_bwd, Ms3);
        else ADM_MS3_LAUNCH(k_ms3_bwd, Ms3p64);
        return ADM_OK;
    }
#define ADM_BWD(CFG, MINB, CT, GCK)                                                                                  \
    do {                                                                                                           \
        const size_t sm_ = MsSmem<CFG>::BYTES;                                                                     \
        if (set_smem(k_multislice_bwd<CFG, MINB, CT, GCK>, sm_)) return ADM_ERR_CUDA;                              \
        int grid_;                                                                                                 \
        if (ms_schedule(p, ms_resident_ctas(k_multislice_bwd<CFG, MINB, CT, GCK>, CFG::NTHREADS, sm_), workspace,  \
                        CFG::P, grid_))                                                                            \
            return ADM_ERR_CUDA;                                                                                   \
        ADM_LAUNCH((k_multislice_bwd<CFG, MINB, CT, GCK>), grid_, CFG::NTHREADS, sm_, stream, map, p);             \
    } while (0)
    if (p.g_ckpt != nullptr) {          // with the adjoint-wavefield checkpoints (packed arithmetic only)
        if (cfg->P == 72) ADM_BWD(Cfg72, 2, cpk, true);
        else ADM_BWD(Cfg64, 2, cpk, true);
        return ADM_OK;
    }
    if (cfg->P == 72) {
        if (var == 4) ADM_BWD(Cfg72, 2, cpx, false);
        else ADM_BWD(Cfg72, 2, cpk, false);
    } else {
        if (var == 4) ADM_BWD(Cfg64, 2, cpx, false);
        else ADM_BWD(Cfg64, 2, cpk, false);
    }
    return ADM_OK;
}
