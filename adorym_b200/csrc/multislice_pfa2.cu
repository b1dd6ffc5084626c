// Two-pass prime-factor multislice kernels for P = 72 (sm_100a): the dominant kernels of config D.
//
// The 72-point axis is indexed by the Good-Thomas map n = (9 a + 8 b) mod 72 (a < 8, b < 9), which turns the 2-D FFT
// of the [72][72] wavefield into a 4-D DFT over (ay, by, ax, bx) with NO twiddle factors, whose four 1-D stages
// commute.  Grouping them as
//
//     R pass  [Bx^-1 By^-1] . ( psi checkpoint, psi (1 + T') ) . [By Bx]      thread (ay, ax)  holds the 9 x 9 block (by, bx)
//     S pass  [Ay Ax] . H . [Ax^-1 Ay^-1]                                     thread (by, bx)  holds the 8 x 8 block (ay, ax)
//
// gives TWO shared-memory round trips of the wavefield per slice instead of the six of the row/column scheme in
// multislice.cu (whose profile showed the shared-memory pipe at 70 % and 6 CTA barriers per slice): all butterflies of
// a pass run in registers (81 or 64 complex values per thread).  Shared-memory layout of the wavefield between the
// passes: element (ay, ax, by, bx) at (8 ay + ax) * 81 + (9 by + bx); the R pass reads with lane stride 81 (odd: no
// bank conflicts for 8-byte accesses), the S pass with lane stride 1.  The T' tile is in natural [y][TP] order (TMA
// box); R-pass lanes (ay, ax) touch x = ax (mod 8) and rows that differ by an odd count, so a row pitch TP = 8 (mod
// 16) complex is conflict free: patches starting at an even column use a 72-wide box (TP = 72), odd ones a 74-wide
// box from the column below (two tensor maps, chosen per CTA).
//
// The psi checkpoint of a slice is stored as [e = 9 by + bx][r = 8 ay + ax] (coalesced 512-byte rows; 5184 complex per
// slice); the adjoint bulk-copies it into the wavefield buffer once every R thread has taken its block into registers,
// overwrites the T' tile in place with dL/dT = g' conj(psi_z) and scatter-adds the tile with one TMA reduce per slice.
//
// Reference behaviour reproduced: as multislice.cu (adorym/propagate.py:208-280, forward_model.py:313-331,
// util.py:1327-1351; adjoint of the same chain instead of the autograd tape, wrappers.py:300-331).
// Restrictions (otherwise multislice.cu serves the call): P = 72, separable in-object transfer function, detector
// transform 'far field' or none.
#include <stdlib.h>
#include "fft_passes.cuh"
#include "tma.cuh"
#include "ms_params.cuh"

namespace ms2 {
constexpr int P = 72;
constexpr int NR = 64;          // R-pass threads (ay, ax)
constexpr int NS = 81;          // S-pass threads (by, bx)
constexpr int NT = 96;          // CTA size
constexpr int DMA = 95;         // the thread that drives tile loads / reduces (idle in both passes)
constexpr int NPSI = P * P;     // 5184 complex per wavefield / checkpoint slice
constexpr int TPMAX = P + 2;
constexpr unsigned PSI_BYTES = NPSI * sizeof(cpx);
constexpr size_t OFF_PSI = sizeof(cpx) * P * TPMAX;            // 42624 (multiple of 128)
constexpr size_t OFF_HS = OFF_PSI + PSI_BYTES;
constexpr size_t OFF_BAR = OFF_HS + sizeof(cpx) * 2 * P;
constexpr size_t SMEM_BYTES = OFF_BAR + 32;

ADM_HDM int pos72(int a, int b) { return (9 * a + 8 * b) % 72; }
// frequency held at spectrum position (k1, k2)
ADM_HDM int freq72(int k1, int k2) { return (9 * k1 + 64 * k2) % 72; }

template <bool INV, int ST, class T> ADM_DEV void dft9s(T* a) {
    T t[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) t[i] = a[i * ST];
    dft9<INV>(t);
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i * ST] = t[i];
}
template <bool INV, int ST, class T> ADM_DEV void dft8s(T* a) {
    T t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = a[i * ST];
    dft8<INV>(t);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i * ST] = t[i];
}

// v[9 by + bx]: 9-point transforms along bx then by (or by then bx; they commute)
template <bool INV, class T> ADM_DEV void dft9x9(T* v) {
#pragma unroll
    for (int by = 0; by < 9; ++by) dft9s<INV, 1>(v + 9 * by);
#pragma unroll
    for (int bx = 0; bx < 9; ++bx) dft9s<INV, 9>(v + bx);
}

ADM_DEV void named_sync_r() {
#ifndef ADM_HOST_EMU
    asm volatile("bar.sync 1, 64;\n" ::: "memory");
#else
    emu_named_sync(1, 64);
#endif
}
ADM_DEV void l2_prefetch(const void* g, unsigned bytes) {
#ifndef ADM_HOST_EMU
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(g), "r"(bytes) : "memory");
#else
    (void)g; (void)bytes;
#endif
}

// S pass: u[8 jy + jx] <- IDFT8x8( H . DFT8x8(u) ) for the thread's (by, bx); hs = [hy | hx] indexed by spectrum position
template <class T> ADM_DEV void s_pass(T* psi, const T* hs, int s) {
    const int by = s / 9, bx = s - 9 * by;
    T u[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) u[j] = psi[j * NS + s];
#pragma unroll
    for (int jx = 0; jx < 8; ++jx) dft8s<false, 8>(u + jx);          // along ay
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const T h = hs[pos72(k, by)];
#pragma unroll
        for (int jx = 0; jx < 8; ++jx) u[8 * k + jx] = cmul(u[8 * k + jx], h);
    }
#pragma unroll
    for (int jy = 0; jy < 8; ++jy) dft8s<false, 1>(u + 8 * jy);      // along ax
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const T h = hs[P + pos72(k, bx)];
#pragma unroll
        for (int jy = 0; jy < 8; ++jy) u[8 * jy + k] = cmul(u[8 * jy + k], h);
    }
#pragma unroll
    for (int jy = 0; jy < 8; ++jy) dft8s<true, 1>(u + 8 * jy);
#pragma unroll
    for (int jx = 0; jx < 8; ++jx) dft8s<true, 8>(u + jx);
#pragma unroll
    for (int j = 0; j < 64; ++j) psi[j * NS + s] = u[j];
}

template <class T> struct Smem {
    T* tile; T* psi; T* hs; adm_mbar_t* barT; adm_mbar_t* barC;
    ADM_HDM explicit Smem(unsigned char* raw) {
        tile = (T*)raw;
        psi = (T*)(raw + OFF_PSI);
        hs = (T*)(raw + OFF_HS);
        barT = (adm_mbar_t*)(raw + OFF_BAR);
        barC = barT + 1;
    }
};

// ================================================================================================
// Forward
// ================================================================================================
template <class T> __global__ void __launch_bounds__(NT, 2) k_ms2_fwd(const ADM_GRID_CONSTANT AdmTileMap map_even,
                                                   const ADM_GRID_CONSTANT AdmTileMap map_odd, MsParams p) {
    ADM_SMEM_DECL
    const Smem<T> sm(smem_raw);
    const int t = threadIdx.x;
    const int wf = blockIdx.x;
    const int b = wf / p.M, m = wf % p.M;
    const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
    const int odd = px & 1, pxa = px & ~1;
    const int TP = P + 2 * odd;
    const AdmTileMap* tmap = odd ? &map_odd : &map_even;
    const unsigned tile_bytes = (unsigned)(sizeof(T) * P * TP);
    const T* probe = (const T*)p.probe + (size_t)(p.probe_per_pattern ? wf : m) * NPSI;
    T* ckpt = (T*)p.ckpt + (size_t)wf * p.Nz * NPSI;
    T* det = (T*)p.det + (size_t)wf * NPSI;
    // R-thread geometry
    const int ay = (t >> 3) & 7, ax = t & 7;

    for (int i = t; i < 2 * P; i += NT) sm.hs[i] = ((const T*)p.Hp)[i];
    if (t == DMA) mbar_init(sm.barT, 1);
    __syncthreads();
    if (t == DMA) {
        mbar_expect_tx(sm.barT, tile_bytes);
        tma_load_tile(sm.tile, tmap, 2 * pxa, py, 0, sm.barT);
    }

    for (int z = 0; z < p.Nz; ++z) {
        const bool last = (z == p.Nz - 1);
        if (t < NR) {
            T v[81];
            if (z == 0) {
#pragma unroll
                for (int by = 0; by < 9; ++by)
#pragma unroll
                    for (int bx = 0; bx < 9; ++bx) v[9 * by + bx] = probe[pos72(ay, by) * P + pos72(ax, bx)];
            } else {
#pragma unroll
                for (int e = 0; e < 81; ++e) v[e] = sm.psi[t * NS + e];
                dft9x9<true>(v);
            }
            mbar_wait(sm.barT, z & 1);
            T* ck = ckpt + (size_t)z * NPSI + t;
#pragma unroll
            for (int by = 0; by < 9; ++by) {
                const T* row = sm.tile + pos72(ay, by) * TP + odd;
#pragma unroll
                for (int bx = 0; bx < 9; ++bx) {
                    const int e = 9 * by + bx;
                    ck[e * NR] = v[e];
                    v[e] = modulate(v[e], row[pos72(ax, bx)]);
                }
            }
            if (last && p.free_prop == ADM_PROP_NONE) {
#pragma unroll
                for (int by = 0; by < 9; ++by)
#pragma unroll
                    for (int bx = 0; bx < 9; ++bx) det[pos72(ay, by) * P + pos72(ax, bx)] = v[9 * by + bx];
            } else {
                if (last && p.ff_conj) {
#pragma unroll
                    for (int e = 0; e < 81; ++e) v[e] = cconj(v[e]);
                }
                dft9x9<false>(v);
#pragma unroll
                for (int e = 0; e < 81; ++e) sm.psi[t * NS + e] = v[e];
            }
        }
        if (last && p.free_prop == ADM_PROP_NONE) break;
        __syncthreads();
        if (t == DMA && !last) {   // every R thread has consumed tile(z): fetch tile(z+1) behind the S pass
            mbar_expect_tx(sm.barT, tile_bytes);
            tma_load_tile(sm.tile, tmap, 2 * pxa, py, z + 1, sm.barT);
        }
        if (last) {
            // far field: the remaining 8-point stages, stored fftshift-ed in natural order
            if (t < NS) {
                const int by = t / 9, bx = t - 9 * by;
                T u[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) u[j] = sm.psi[j * NS + t];
#pragma unroll
                for (int jx = 0; jx < 8; ++jx) dft8s<false, 8>(u + jx);
#pragma unroll
                for (int jy = 0; jy < 8; ++jy) dft8s<false, 1>(u + 8 * jy);
#pragma unroll
                for (int ky = 0; ky < 8; ++ky) {
                    const int fy = (freq72(ky, by) + P / 2) % P;
#pragma unroll
                    for (int kx = 0; kx < 8; ++kx) {
                        const int fx = (freq72(kx, bx) + P / 2) % P;
                        const T o = cscale(u[8 * ky + kx], p.ff_scale);
                        det[fy * P + fx] = p.ff_conj ? cconj(o) : o;
                    }
                }
            }
            break;
        }
        if (t < NS) s_pass(sm.psi, sm.hs, t);
        __syncthreads();
    }
}

// ================================================================================================
// Adjoint
// ================================================================================================
template <class T> __global__ void __launch_bounds__(NT, 2) k_ms2_bwd(const ADM_GRID_CONSTANT AdmTileMap tmap_even,
                                                   const ADM_GRID_CONSTANT AdmTileMap tmap_odd,
                                                   const ADM_GRID_CONSTANT AdmTileMap gmap_even,
                                                   const ADM_GRID_CONSTANT AdmTileMap gmap_odd, MsParams p) {
    ADM_SMEM_DECL
    const Smem<T> sm(smem_raw);
    const int t = threadIdx.x;
    const int wf = blockIdx.x;
    const int b = wf / p.M, m = wf % p.M;
    const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
    const int odd = px & 1, pxa = px & ~1;
    const int TP = P + 2 * odd;
    const AdmTileMap* tmap = odd ? &tmap_odd : &tmap_even;
    const AdmTileMap* gmap = odd ? &gmap_odd : &gmap_even;
    const unsigned tile_bytes = (unsigned)(sizeof(T) * P * TP);
    const T* ckpt = (T*)p.ckpt + (size_t)wf * p.Nz * NPSI;
    const T* gdet = (const T*)p.det + (size_t)wf * NPSI;
    const int ay = (t >> 3) & 7, ax = t & 7;
    const bool touches = (pxa + P >= 0) && (py + P >= 0);   // (a patch entirely outside contributes nothing)

    for (int i = t; i < 2 * P; i += NT) sm.hs[i] = ((const T*)p.Hp)[i];
    if (t == DMA) { mbar_init(sm.barT, 1); mbar_init(sm.barC, 1); }
    __syncthreads();
    if (t == DMA) {
        mbar_expect_tx(sm.barT, tile_bytes);
        tma_load_tile(sm.tile, tmap, 2 * pxa, py, p.Nz - 1, sm.barT);
        l2_prefetch(ckpt + (size_t)(p.Nz - 1) * NPSI, PSI_BYTES);
    }

    // ---- adjoint of the far-field 8-point stages --------------------------------------------------------------
    if (p.free_prop == ADM_PROP_FAR) {
        if (t < NS) {
            const int by = t / 9, bx = t - 9 * by;
            T u[64];
#pragma unroll
            for (int ky = 0; ky < 8; ++ky) {
                const int fy = (freq72(ky, by) + P / 2) % P;
#pragma unroll
                for (int kx = 0; kx < 8; ++kx) {
                    const int fx = (freq72(kx, bx) + P / 2) % P;
                    const T o = cscale(gdet[fy * P + fx], p.ff_scale);
                    u[8 * ky + kx] = p.ff_conj ? cconj(o) : o;
                }
            }
#pragma unroll
            for (int jy = 0; jy < 8; ++jy) dft8s<true, 1>(u + 8 * jy);
#pragma unroll
            for (int jx = 0; jx < 8; ++jx) dft8s<true, 8>(u + jx);
#pragma unroll
            for (int j = 0; j < 64; ++j) sm.psi[j * NS + t] = u[j];
        }
        __syncthreads();
    }

    int it = 0;
    for (int z = p.Nz - 1; z >= 0; --z, ++it) {
        const bool last = (z == p.Nz - 1);
        if (t < NR) {
            T v[81];
            const bool from_det = last && p.free_prop == ADM_PROP_NONE;
            if (from_det) {
#pragma unroll
                for (int by = 0; by < 9; ++by)
#pragma unroll
                    for (int bx = 0; bx < 9; ++bx) v[9 * by + bx] = gdet[pos72(ay, by) * P + pos72(ax, bx)];
            } else {
#pragma unroll
                for (int e = 0; e < 81; ++e) v[e] = sm.psi[t * NS + e];
            }
            named_sync_r();     // every R thread holds its block: the wavefield buffer is free for the checkpoint
            if (t == 0) {
                mbar_expect_tx(sm.barC, PSI_BYTES);
                bulk_load(sm.psi, ckpt + (size_t)z * NPSI, PSI_BYTES, sm.barC);
                if (z > 0) l2_prefetch(ckpt + (size_t)(z - 1) * NPSI, PSI_BYTES);
            }
            if (!from_det) {
                dft9x9<true>(v);
                if (last && p.ff_conj) {
#pragma unroll
                    for (int e = 0; e < 81; ++e) v[e] = cconj(v[e]);
                }
            }
            mbar_wait(sm.barT, it & 1);
            if (odd) {   // the two pad columns of the 74-wide tile hold neighbouring T' values: nothing to add there
                for (int i = t; i < 2 * P; i += NR) sm.tile[(i >> 1) * TP + ((i & 1) ? P + 1 : 0)] = tmk<T>(0.f, 0.f);
            }
            mbar_wait(sm.barC, it & 1);
#pragma unroll
            for (int by = 0; by < 9; ++by) {
                T* row = sm.tile + pos72(ay, by) * TP + odd;
#pragma unroll
                for (int bx = 0; bx < 9; ++bx) {
                    const int e = 9 * by + bx;
                    const T w = sm.psi[e * NR + t];        // psi_z (wave entering the slice)
                    const int x = pos72(ax, bx);
                    const T tp = row[x];
                    row[x] = cmulc(v[e], w);                 // dL/dT = g' conj(psi_z)
                    v[e] = modulate_conj(v[e], tp);          // g_psi = g' conj(1 + T')
                }
            }
            fence_proxy_async();      // the dL/dT tile is about to be read by the TMA unit
            if (z > 0) {
                dft9x9<false>(v);
                named_sync_r();       // every R thread has read the checkpoint: the buffer takes the wavefield again
#pragma unroll
                for (int e = 0; e < 81; ++e) sm.psi[t * NS + e] = v[e];
            } else if (p.g_probe != nullptr) {
#pragma unroll
                for (int by = 0; by < 9; ++by)
#pragma unroll
                    for (int bx = 0; bx < 9; ++bx) {
                        const int o = pos72(ay, by) * P + pos72(ax, bx);
                        const T g = v[9 * by + bx];
                        if (p.probe_per_pattern) ((T*)p.g_probe)[(size_t)wf * NPSI + o] = g;
                        else red_add2(&p.g_probe[(size_t)m * NPSI + o], cx(g), cy(g));
                    }
            }
        }
        __syncthreads();
        if (t == DMA) {
            if (touches) {
                tma_reduce_add_tile(gmap, 2 * (pxa + P), py + P, z, sm.tile);   // margined volume: never negative
                bulk_commit();
            }
            if (z > 0) {
                bulk_wait_read_all();     // the reduce has drained the tile: fetch T' of slice z-1 behind the S pass
                mbar_expect_tx(sm.barT, tile_bytes);
                tma_load_tile(sm.tile, tmap, 2 * pxa, py, z - 1, sm.barT);
            }
        }
        if (z == 0) break;
        if (t < NS) s_pass(sm.psi, sm.hs, t);
        __syncthreads();
    }
    if (t == DMA) bulk_wait_all();      // shared memory must stay valid until the last reduce has read it
}
}   // namespace ms2

// ================================================================================================
// host side
// ================================================================================================
// Opt-in (ADM_MS_PFA2=1): on B200 the two-pass kernels are correct (parity suite green) but, at 6 warps per SM, slower
// than the six-pass kernels of multislice.cu (profiles/r01e_*): 6.4 / 7.3 ms against 4.0 / 4.9 ms at config D.
int adm_ms2_supported(const adm_multislice_cfg* c) {
    const char* e = getenv("ADM_MS_PFA2");
    if (!e || !atoi(e)) return 0;
    return c->P == 72 && c->h_separable && c->n_slices > 1 && c->free_prop != ADM_PROP_NEAR;
}

// (instantiated on the scalar complex type: with packed register pairs the 81-value blocks spill more, 8.3 / 7.1 ms)
int adm_ms2_fwd_launch(const adm_multislice_cfg* c, const MsParams& p, const float* tvol, void* stream) {
    AdmTileMap me, mo;
    if (adm_make_tile_map(&me, tvol, p.Nz, p.Y, p.X, p.Xp, ms2::P, ms2::P)) return ADM_ERR_CUDA;
    if (adm_make_tile_map(&mo, tvol, p.Nz, p.Y, p.X, p.Xp, ms2::P, ms2::P + 2)) return ADM_ERR_CUDA;
    if (set_smem(ms2::k_ms2_fwd<cpx>, ms2::SMEM_BYTES)) return ADM_ERR_CUDA;
    ADM_LAUNCH(ms2::k_ms2_fwd<cpx>, c->B * c->M, ms2::NT, ms2::SMEM_BYTES, stream, me, mo, p);
    return ADM_OK;
}

int adm_ms2_bwd_launch(const adm_multislice_cfg* c, const MsParams& p, const float* tvol, float* g_tvol, void* stream) {
    AdmTileMap te, to, ge, go;
    if (adm_make_tile_map(&te, tvol, p.Nz, p.Y, p.X, p.Xp, ms2::P, ms2::P)) return ADM_ERR_CUDA;
    if (adm_make_tile_map(&to, tvol, p.Nz, p.Y, p.X, p.Xp, ms2::P, ms2::P + 2)) return ADM_ERR_CUDA;
    const int rows = ADM_GVOL_ROWS(p.Y, ms2::P), gx = p.X + ADM_GVOL_MARGIN(ms2::P), gp = ADM_GVOL_PITCH(p.X, ms2::P);
    if (adm_make_tile_map(&ge, g_tvol, p.Nz, rows, gx, gp, ms2::P, ms2::P)) return ADM_ERR_CUDA;
    if (adm_make_tile_map(&go, g_tvol, p.Nz, rows, gx, gp, ms2::P, ms2::P + 2)) return ADM_ERR_CUDA;
    if (set_smem(ms2::k_ms2_bwd<cpx>, ms2::SMEM_BYTES)) return ADM_ERR_CUDA;
    ADM_LAUNCH(ms2::k_ms2_bwd<cpx>, c->B * c->M, ms2::NT, ms2::SMEM_BYTES, stream, te, to, ge, go, p);
    return ADM_OK;
}
