// Register-resident line FFTs for P = 72 (sm_100a): TWO shared-memory round trips of the wavefield per slice.
//
// A thread owns one whole line of the [72][72] wavefield (72 packed complex values = 144 registers) and performs the
// complete 72-point prime-factor FFT of that line in registers (9 radix-8 + 8 radix-9 packed butterflies, static
// indices n = (9 a + 8 b) mod 72, no twiddles), so a slice is
//
//     pass Y  (thread = column)  [FFTy^-1] . psi checkpoint . psi (1 + T') . [FFTy]
//     pass X  (thread = row)     [FFTx] . H . [FFTx^-1]
//
// with 2 CTA barriers, against 6 passes / 6 barriers in multislice.cu, whose profile (profiles/r01l) shows the
// shared-memory pipe as the bound once the arithmetic is packed.  Layouts are the natural ones of multislice.cu
// (wavefield [y][73], T' tile and checkpoint [y][74] with the px & 1 column offset), so the data movement is
// unchanged: T' tile by TMA prefetched one slice ahead, checkpoint by cp.async.bulk, dL/dT tile scatter-added by one
// TMA reduce per slice, and the spectrum order / transfer-function tables are those of multislice.cu.
// 72 lines -> 3 warps of 24 active lanes; 86 KB of shared memory, 2 CTAs per SM.
//
// Restrictions (otherwise multislice.cu serves the call): P = 72, separable in-object transfer function, detector
// transform 'far field' or none.  Reference behaviour reproduced: as multislice.cu.
#include <stdlib.h>
#include "fft_passes.cuh"
#include "tma.cuh"
#include "ms_params.cuh"

namespace msc {
constexpr int P = 72, S = 73, TP = P + 2;
constexpr int NT = 96;           // 3 warps x (24 lines + 8 idle lanes)
constexpr int DMA = 95;          // an idle lane drives the bulk copies
constexpr size_t TILE_BYTES = sizeof(cpx) * P * TP;                    // 42624 (multiple of 128)
constexpr size_t OFF_PSI = TILE_BYTES;
constexpr size_t OFF_HS = OFF_PSI + sizeof(cpx) * P * S;
constexpr size_t OFF_BAR = OFF_HS + sizeof(cpx) * 2 * P;
constexpr size_t SMEM_BYTES = OFF_BAR + 16;

ADM_HDM constexpr int pos72(int a, int b) { return (9 * a + 8 * b) % 72; }
ADM_HDM constexpr int freq72(int p) { return (9 * (p % 8) + 64 * ((9 - p % 9) % 9)) % 72; }   // = perm_freq<Cfg72>

// 72-point DFT of a line held in registers, in place: v[n] (natural position) -> spectrum in the position order of
// multislice.cu (position (9 k1 + 8 k2) mod 72 holds frequency (9 k1 + 64 k2) mod 72); INV = conjugate, unnormalised
template <bool INV, class T> ADM_DEV void fft72(T* v) {
#pragma unroll
    for (int b = 0; b < 9; ++b) {
        T t[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) t[a] = v[pos72(a, b)];
        dft8<INV>(t);
#pragma unroll
        for (int a = 0; a < 8; ++a) v[pos72(a, b)] = t[a];
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        T t[9];
#pragma unroll
        for (int b = 0; b < 9; ++b) t[b] = v[pos72(a, b)];
        dft9<INV>(t);
#pragma unroll
        for (int b = 0; b < 9; ++b) v[pos72(a, b)] = t[b];
    }
}

// pass X: row l <- IFFTx( H . FFTx(row l) ), hs = [hy | hx] by spectrum position (hx carries 1 / P^2)
template <class T> ADM_DEV void pass_x(T* psi, const T* hs, int l) {
    T v[72];
#pragma unroll
    for (int x = 0; x < 72; ++x) v[x] = psi[l * S + x];
    fft72<false>(v);
    const T hyl = hs[l];
#pragma unroll
    for (int x = 0; x < 72; ++x) v[x] = cmul(v[x], cmul(hs[P + x], hyl));
    fft72<true>(v);
#pragma unroll
    for (int x = 0; x < 72; ++x) psi[l * S + x] = v[x];
}

template <class T> struct Sm {
    T* tile; T* psi; T* hs; adm_mbar_t* bar;
    ADM_HDM explicit Sm(unsigned char* raw) {
        tile = (T*)raw; psi = (T*)(raw + OFF_PSI); hs = (T*)(raw + OFF_HS); bar = (adm_mbar_t*)(raw + OFF_BAR);
    }
};
ADM_DEV int line_of(int t) { const int i = t & 31; return i < 24 ? (t >> 5) * 24 + i : -1; }

// ================================================================================================
// Forward
// ================================================================================================
// The slice loop runs FOUR line transforms per slice through ONE copy of the 72-point FFT code (a rolled loop over
// `step`), because the first version -- forward and inverse transforms unrolled in both passes, 3.7 k SASS per slice --
// spent 37 % of its issue slots waiting for instruction fetch (profiles/r01m).  The inverse transform is obtained from
// the forward code by conjugation, IFFT(x) = conj(FFT(conj(x))), and the conjugations are folded into neighbours:
//   step 0  column <- psi (holds conj of the y spectrum)   FFT -> conj(psi_z): conjugate, checkpoint, modulate
//   step 1  (registers)                                     FFT -> y spectrum, store column            | barrier
//   step 2  row <- psi                                      FFT -> x spectrum: multiply by H, conjugate
//   step 3  (registers)                                     FFT -> conj(propagated y spectrum), store | barrier
template <class T> __global__ void __launch_bounds__(NT, 2) k_msc_fwd(const ADM_GRID_CONSTANT AdmTileMap tmap, MsParams p) {
    ADM_SMEM_DECL
    const Sm<T> sm(smem_raw);
    const int t = threadIdx.x, l = line_of(t);
    const bool active = l >= 0;
    const int wf = blockIdx.x;
    const int b = wf / p.M, m = wf % p.M;
    const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
    const int pxa = px & ~1, lt = l + (px & 1);
    const T* probe = (const T*)p.probe + (size_t)(p.probe_per_pattern ? wf : m) * P * P;
    T* ckpt = (T*)p.ckpt + (size_t)wf * p.Nz * P * TP;
    T* det = (T*)p.det + (size_t)wf * P * P;

    for (int i = t; i < 2 * P; i += NT) sm.hs[i] = ((const T*)p.Hp)[i];
    if (t == DMA) mbar_init(sm.bar, 1);
    __syncthreads();
    if (t == DMA) {
        mbar_expect_tx(sm.bar, (unsigned)TILE_BYTES);
        tma_load_tile(sm.tile, &tmap, 2 * pxa, py, 0, sm.bar);
    }
    // prologue: the loop's step 0 expects conj(y spectrum) of the wave entering the slice
    if (active) {
        T v[72];
#pragma unroll
        for (int y = 0; y < 72; ++y) v[y] = probe[y * P + l];
        fft72<false>(v);
        const float s = 1.0f / (float)P;            // step 0 undoes it with an unnormalised transform
#pragma unroll
        for (int y = 0; y < 72; ++y) sm.psi[y * S + l] = cscale(cconj(v[y]), s);
    }
    __syncthreads();

    T v[72];
    for (int z = 0; z < p.Nz; ++z) {
        const bool last = (z == p.Nz - 1);
        const bool exit_wave = last && p.free_prop == ADM_PROP_NONE;
#pragma unroll 1
        for (int step = 0; step < 4; ++step) {
            if (last && step == 2) break;
            if (active) {
                if (step == 0) {
#pragma unroll
                    for (int y = 0; y < 72; ++y) v[y] = sm.psi[y * S + l];
                } else if (step == 2) {
#pragma unroll
                    for (int x = 0; x < 72; ++x) v[x] = sm.psi[l * S + x];
                }
                if (!(exit_wave && step == 1)) fft72<false>(v);
                if (step == 0) {
                    mbar_wait(sm.bar, z & 1);
                    T* ck = ckpt + (size_t)z * P * TP + lt;
#pragma unroll
                    for (int y = 0; y < 72; ++y) {
                        const T w = cconj(v[y]);            // psi_z
                        ck[y * TP] = w;
                        v[y] = modulate(w, sm.tile[y * TP + lt]);
                        if (last && p.ff_conj) v[y] = cconj(v[y]);
                    }
                } else if (step == 1) {
                    if (exit_wave) {
#pragma unroll
                        for (int y = 0; y < 72; ++y) det[y * P + l] = v[y];
                    } else {
#pragma unroll
                        for (int y = 0; y < 72; ++y) sm.psi[y * S + l] = v[y];
                    }
                } else if (step == 2) {
                    const T hyl = sm.hs[l];
#pragma unroll
                    for (int x = 0; x < 72; ++x) v[x] = cconj(cmul(v[x], cmul(sm.hs[P + x], hyl)));
                } else {
#pragma unroll
                    for (int x = 0; x < 72; ++x) sm.psi[l * S + x] = v[x];
                }
            }
            if (step & 1) {
                __syncthreads();
                if (step == 1 && t == DMA && !last) {   // every line has consumed tile(z): fetch tile(z+1)
                    mbar_expect_tx(sm.bar, (unsigned)TILE_BYTES);
                    tma_load_tile(sm.tile, &tmap, 2 * pxa, py, z + 1, sm.bar);
                }
            }
        }
    }
    // far field: forward FFT along x of row l (spectrum position l along y), stored fftshift-ed
    if (p.free_prop == ADM_PROP_FAR && active) {
#pragma unroll
        for (int x = 0; x < 72; ++x) v[x] = sm.psi[l * S + x];
        fft72<false>(v);
        const int fy = (freq72(l) + P / 2) % P;
#pragma unroll
        for (int x = 0; x < 72; ++x) {
            const int fx = (freq72(x) + P / 2) % P;
            const T o = cscale(v[x], p.ff_scale);
            det[fy * P + fx] = p.ff_conj ? cconj(o) : o;
        }
    }
}

// ================================================================================================
// Adjoint
// ================================================================================================
template <class T> __global__ void __launch_bounds__(NT, 2) k_msc_bwd(const ADM_GRID_CONSTANT AdmTileMap gmap, MsParams p) {
    ADM_SMEM_DECL
    const Sm<T> sm(smem_raw);
    T* ckb = sm.tile;            // psi checkpoint of the slice (bulk copy); overwritten in place by dL/dT
    const int t = threadIdx.x, l = line_of(t);
    const bool active = l >= 0;
    const int wf = blockIdx.x;
    const int b = wf / p.M, m = wf % p.M;
    const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
    const int pxa = px & ~1, lt = l + (px & 1);
    const int gx = px + l;
    const bool x_in = active && (gx >= 0) && (gx < p.X);
    const T* ckpt = (const T*)p.ckpt + (size_t)wf * p.Nz * P * TP;
    const T* gdet = (const T*)p.det + (size_t)wf * P * P;
    const size_t slice_stride = (size_t)p.Y * p.Xp;
    const T* tcol = (const T*)p.tvol + (x_in ? gx : 0);
    const bool touches = (pxa + P >= 0) && (py + P >= 0);

    for (int i = t; i < 2 * P; i += NT) sm.hs[i] = ((const T*)p.Hp)[i];
    if (t == DMA) mbar_init(sm.bar, 1);
    __syncthreads();
    if (t == DMA) {   // prefetch the checkpoint of the last slice
        mbar_expect_tx(sm.bar, (unsigned)TILE_BYTES);
        bulk_load(ckb, ckpt + (size_t)(p.Nz - 1) * P * TP, (unsigned)TILE_BYTES, sm.bar);
    }
    if (p.free_prop == ADM_PROP_FAR) {
        // adjoint of the far-field FFT along x
        if (active) {
            T v[72];
            const int fy = (freq72(l) + P / 2) % P;
#pragma unroll
            for (int x = 0; x < 72; ++x) {
                const int fx = (freq72(x) + P / 2) % P;
                const T o = cscale(gdet[fy * P + fx], p.ff_scale);
                v[x] = p.ff_conj ? cconj(o) : o;
            }
            fft72<true>(v);
#pragma unroll
            for (int x = 0; x < 72; ++x) sm.psi[l * S + x] = v[x];
        }
        __syncthreads();
    }

    int phase = 0;
    for (int z = p.Nz - 1; z >= 0; --z, phase ^= 1) {
        const bool last = (z == p.Nz - 1);
        const bool from_det = last && p.free_prop == ADM_PROP_NONE;
        const T* tz = tcol + (size_t)z * slice_stride;
        // ---- pass Y: [FFTy^-1] . dL/dT = g' conj(psi_z) . g = g' conj(1 + T') . [FFTy] ---------------------------
        if (active) {
            T v[72];
            if (from_det) {
#pragma unroll
                for (int y = 0; y < 72; ++y) v[y] = gdet[y * P + l];
            } else {
#pragma unroll
                for (int y = 0; y < 72; ++y) v[y] = sm.psi[y * S + l];
                fft72<true>(v);
                if (last && p.ff_conj) {
#pragma unroll
                    for (int y = 0; y < 72; ++y) v[y] = cconj(v[y]);
                }
            }
            mbar_wait(sm.bar, phase);
            // the two pad columns of the tile are never written by the forward pass: add zeros there
            if (l < 2 * 24) {
                for (int i = l; i < 2 * P; i += 48)
                    ckb[(i >> 1) * TP + ((i & 1) ? P + 1 : ((px & 1) ? 0 : P))] = tmk<T>(0.f, 0.f);
            }
#pragma unroll
            for (int y = 0; y < 72; ++y) {
                const int gy = py + y;
                const bool inb = x_in && (unsigned)gy < (unsigned)p.Y;
                const T tp = inb ? tz[(unsigned)(gy * p.Xp)] : tmk<T>(0.f, 0.f);
                const T w = ckb[y * TP + lt];               // psi_z (wave entering the slice)
                ckb[y * TP + lt] = cmulc(v[y], w);          // dL/dT = g' conj(psi_z)
                v[y] = modulate_conj(v[y], tp);             // g_psi = g' conj(1 + T')
            }
            if (z > 0) {
                fft72<false>(v);
#pragma unroll
                for (int y = 0; y < 72; ++y) sm.psi[y * S + l] = v[y];
            } else if (p.g_probe != nullptr) {
#pragma unroll
                for (int y = 0; y < 72; ++y) {
                    if (p.probe_per_pattern) ((T*)p.g_probe)[(size_t)wf * P * P + y * P + l] = v[y];
                    else red_add2(&p.g_probe[(size_t)m * P * P + y * P + l], cx(v[y]), cy(v[y]));
                }
            }
            fence_proxy_async();      // the dL/dT tile is about to be read by the TMA unit
        }
        __syncthreads();
        if (t == DMA) {
            if (touches) {
                tma_reduce_add_tile(&gmap, 2 * (pxa + P), py + P, z, ckb);   // margined volume: never negative
                bulk_commit();
            }
            if (z > 0) {   // once the reduce has drained the tile: checkpoint of slice z-1, behind pass X
                bulk_wait_read_all();
                mbar_expect_tx(sm.bar, (unsigned)TILE_BYTES);
                bulk_load(ckb, ckpt + (size_t)(z - 1) * P * TP, (unsigned)TILE_BYTES, sm.bar);
            }
        }
        if (z == 0) break;
        if (active) pass_x(sm.psi, sm.hs, l);
        __syncthreads();
    }
    if (t == DMA) bulk_wait_all();      // shared memory must stay valid until the last reduce has read it
}
}   // namespace msc

// ================================================================================================
// host side
// ================================================================================================
// Opt-in (ADM_MS_COL=1).  Measured on B200 at config D (profiles/r01m, r01n): forward 4.30 ms with all transforms
// unrolled (37 % of the issue slots lost to instruction fetch), 4.12 ms with the rolled loop above (instruction fetch
// gone, fixed-latency dependency stalls 35 % at 6 warps per SM), against 3.60 ms for the six-pass kernel; the adjoint
// (11.8 ms) still loads T' per element from global memory and waits on it (long scoreboard 43 %).
int adm_msc_supported(const adm_multislice_cfg* c) {
    const char* e = getenv("ADM_MS_COL");
    if (!e || !atoi(e)) return 0;
    return c->P == 72 && c->h_separable && c->n_slices > 1 && c->free_prop != ADM_PROP_NEAR;
}

int adm_msc_fwd_launch(const adm_multislice_cfg* c, const MsParams& p, const float* tvol, void* stream) {
    AdmTileMap map;
    if (adm_make_tile_map(&map, tvol, p.Nz, p.Y, p.X, p.Xp, msc::P, msc::TP)) return ADM_ERR_CUDA;
    const char* e = getenv("ADM_MSC_SCALAR");
    if (e && atoi(e)) {      // scalar instantiation for A/B runs
        if (set_smem(msc::k_msc_fwd<cpx>, msc::SMEM_BYTES)) return ADM_ERR_CUDA;
        ADM_LAUNCH(msc::k_msc_fwd<cpx>, c->B * c->M, msc::NT, msc::SMEM_BYTES, stream, map, p);
        return ADM_OK;
    }
    if (set_smem(msc::k_msc_fwd<cpk>, msc::SMEM_BYTES)) return ADM_ERR_CUDA;
    ADM_LAUNCH(msc::k_msc_fwd<cpk>, c->B * c->M, msc::NT, msc::SMEM_BYTES, stream, map, p);
    return ADM_OK;
}

int adm_msc_bwd_launch(const adm_multislice_cfg* c, const MsParams& p, float* g_tvol, void* stream) {
    AdmTileMap map;
    if (adm_make_tile_map(&map, g_tvol, p.Nz, ADM_GVOL_ROWS(p.Y, msc::P), p.X + ADM_GVOL_MARGIN(msc::P),
                          ADM_GVOL_PITCH(p.X, msc::P), msc::P, msc::TP))
        return ADM_ERR_CUDA;
    if (set_smem(msc::k_msc_bwd<cpk>, msc::SMEM_BYTES)) return ADM_ERR_CUDA;
    ADM_LAUNCH(msc::k_msc_bwd<cpk>, c->B * c->M, msc::NT, msc::SMEM_BYTES, stream, map, p);
    return ADM_OK;
}
