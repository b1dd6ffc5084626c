// Large-frame near-field path (sm_100a): full-field holography frames of 128..2048 pixels per side, one thin object
// slice, several propagation distances (MultiDistModel, adorym/forward_model.py:819-1092) and the sub-pixel Fourier
// shift of measured frames (util.py:380-397 as used at forward_model.py:1079-1088).
//
// A frame does not fit one CTA's shared memory, so a 2-D FFT is two passes over an L2-resident array (8 MB at
// 1024^2, against 126 MB of L2): a ROW kernel (a few rows per CTA) and a COLUMN kernel (a tile of 4096/N columns
// per CTA).  Everything between two transposes is fused into the pass that has the data in shared memory:
//
//   forward   R1  psi0 = probe_m . T(obj)  -> row DIF FFT                                   (obj, probe -> W1)
//             C1  column DIF FFT once, then per distance d: . H_d -> column inverse FFT     (W1 -> det[d])
//             R2  row inverse FFT, in place                                                 (det)
//   adjoint   R2' row forward FFT of dL/d(det)                                              (g_det -> W2)
//             C1' per distance: column forward FFT, . conj(H_d), summed over d in shared memory; one column inverse
//             R1' row inverse FFT, g_probe_m = g conj(T), g_T = sum_m g conj(probe_m), chain rule to (delta, beta)
//
// i.e. FFT2(psi0) is computed once for all distances and the adjoint accumulates over distances in Fourier space:
// 2 + 2 D one-dimensional passes each way instead of the reference's 4 D (propagate.py:271-280 per distance).
//
// The 1-D transform is an in-place decimation-in-frequency radix-8 (+ one radix-2/4) Cooley-Tukey pass sequence; its
// inverse is the conjugate transpose, pass by pass in reverse order (decimation in time), so no reordering pass
// exists: the spectrum lives in digit-reversed order and the transfer functions are permuted once when the plan is
// made (adm_lf_spectrum_order gives the map).  Lines sit in shared memory with a skew of one element per eight
// (lf_pad) and a line pitch = 2 mod 16, which keeps the strided butterfly accesses and the transposing tile
// loads/stores of the column kernel free of bank conflicts.
#include "fft_passes.cuh"
#include "../../include/adorym_b200.h"

#define LF_NT 256          // threads per CTA
#define LF_ROW_ELEMS 2048  // complex elements per CTA in the row kernels  (RPC = 2048 / N rows)
#define LF_COL_ELEMS 4096  // complex elements per CTA in the column kernels (C = 4096 / N columns)

ADM_HD int lf_pad(int n) { return n + (n >> 3); }
ADM_HD int lf_ld(int N) { int l = N + (N >> 3); return l + ((18 - (l & 15)) & 15); }

ADM_DEV void lf_red2(cpx* addr, float a, float b) {
#ifndef ADM_HOST_EMU
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a), "f"(b) : "memory");
#else
    atomicAdd(&addr->x, a);
    atomicAdd(&addr->y, b);
#endif
}
// psi (1 + T') and g conj(1 + T') (T' = T - 1 is what the transmission volume stores)
ADM_DEV cpx modulate(cpx v, cpx tp) {
    return mk(fmaf(-v.y, tp.y, fmaf(v.x, tp.x, v.x)), fmaf(v.y, tp.x, fmaf(v.x, tp.y, v.y)));
}
ADM_DEV cpx modulate_conj(cpx v, cpx tp) {
    return mk(fmaf(v.y, tp.y, fmaf(v.x, tp.x, v.x)), fmaf(v.y, tp.x, fmaf(-v.x, tp.y, v.y)));
}

template <int R, bool INV> ADM_DEV void lf_dft(cpx* v) {
    if constexpr (R == 8) dft8<INV>(v);
    else if constexpr (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else { cpx a = v[0] + v[1], b = v[0] - v[1]; v[0] = a; v[1] = b; }
}

// One pass over nb lines of length N (line f at s + f * ld): blocks of length L, radix R.
//   forward (DIF): butterfly over the R elements j + m L/R, then output k times W_L^{jk}
//   inverse (DIT): the conjugate transpose -- conjugate twiddle first, then the inverse butterfly
template <int N, int L, int R, bool INV> ADM_DEV void lf_pass(cpx* s, int ld, int nb, const cpx* tw, int tid) {
    constexpr int SUB = L / R, PER = N / R, TWS = N / L;
    const int total = nb * PER;
    for (int w = tid; w < total; w += LF_NT) {
        const int f = w / PER, r = w % PER;
        const int blk = r / SUB, j = r % SUB;
        cpx* base = s + f * ld;
        const int i0 = blk * L + j;
        cpx v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = base[lf_pad(i0 + m * SUB)];
        if constexpr (INV && SUB > 1) {
#pragma unroll
            for (int k = 1; k < R; ++k) v[k] = cmulc(v[k], tw[TWS * j * k]);
        }
        lf_dft<R, INV>(v);
        if constexpr (!INV && SUB > 1) {
#pragma unroll
            for (int k = 1; k < R; ++k) v[k] = cmul(v[k], tw[TWS * j * k]);
        }
#pragma unroll
        for (int m = 0; m < R; ++m) base[lf_pad(i0 + m * SUB)] = v[m];
    }
}

// forward: L = N, N/8, ... ; every pass is followed by a barrier
template <int N, int L> ADM_DEV void lf_fwd_from(cpx* s, int ld, int nb, const cpx* tw, int tid) {
    if constexpr (L >= 8) {
        lf_pass<N, L, 8, false>(s, ld, nb, tw, tid);
        __syncthreads();
        lf_fwd_from<N, L / 8>(s, ld, nb, tw, tid);
    } else if constexpr (L > 1) {
        lf_pass<N, L, L, false>(s, ld, nb, tw, tid);
        __syncthreads();
    }
}
// inverse: the same passes in reverse order
template <int N, int L> ADM_DEV void lf_inv_from(cpx* s, int ld, int nb, const cpx* tw, int tid) {
    if constexpr (L >= 8) {
        lf_inv_from<N, L / 8>(s, ld, nb, tw, tid);
        lf_pass<N, L, 8, true>(s, ld, nb, tw, tid);
        __syncthreads();
    } else if constexpr (L > 1) {
        lf_pass<N, L, L, true>(s, ld, nb, tw, tid);
        __syncthreads();
    }
}

template <int N> ADM_DEV void lf_init_twiddles(cpx* tw, int tid) {
    for (int t = tid; t < N; t += LF_NT) {
#ifndef ADM_HOST_EMU
        float sn, cs;
        sincospif(2.0f * (float)t / (float)N, &sn, &cs);
        tw[t] = mk(cs, -sn);
#else
        double a = 2.0 * 3.14159265358979323846 * t / N;
        tw[t] = mk((float)cos(a), (float)-sin(a));
#endif
    }
}

// frequency index (0..N-1, numpy fftfreq order) held at spectrum position p
static int lf_freq_of_pos(int N, int p) {
    int freq = 0, mult = 1, L = N;
    while (L > 1) {
        int R = L >= 8 ? 8 : L, sub = L / R, d = p / sub;
        p %= sub;
        freq += d * mult;
        mult *= R;
        L = sub;
    }
    return freq;
}

// transmission of the thin object (propagate.py:236-246)
ADM_DEV cpx lf_transmission(int unknown_type, float k1, float sgn, cpx o) {
    if (unknown_type == ADM_REAL_IMAG) return o;
    float e = expf(-k1 * o.y), sn, cs;
    adm_sincos(-sgn * k1 * o.x, &sn, &cs);
    return mk(e * cs, e * sn);
}

struct LfParams {
    const cpx* obj;     // [N][N] (delta, beta) | (re, im)
    const cpx* probe;   // [M][N][N]
    const cpx* H;       // [D][N][N] prepared (permuted, scaled; conjugated for the adjoint)
    const cpx* in;      // generic input
    cpx* out;           // generic output
    cpx* g_obj;         // [N][N]
    cpx* g_probe;       // [M][N][N] or null
    int M, D, unknown_type;
    float k1, sgn;
};

// ---- R1: psi0 = probe_m . T -> row DIF FFT -> W1[m] --------------------------------------------------------
template <int N> __global__ void __launch_bounds__(LF_NT) k_lf_rows_mod_fwd(LfParams p) {
    ADM_SMEM_DECL
    constexpr int RPC = LF_ROW_ELEMS / N, EPT = LF_ROW_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* s = (cpx*)smem_raw;
    cpx* tw = s + RPC * ld;
    lf_init_twiddles<N>(tw, tid);
    const int y0 = blockIdx.x * RPC;
    cpx tv[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        tv[k] = lf_transmission(p.unknown_type, p.k1, p.sgn, p.obj[(size_t)(y0 + f) * N + x]);
    }
    for (int m = 0; m < p.M; ++m) {
        const cpx* pr = p.probe + (size_t)m * N * N;
        cpx* o = p.out + (size_t)m * N * N;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
            s[f * ld + lf_pad(x)] = cmul(pr[(size_t)(y0 + f) * N + x], tv[k]);
        }
        __syncthreads();
        lf_fwd_from<N, N>(s, ld, RPC, tw, tid);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
            o[(size_t)(y0 + f) * N + x] = s[f * ld + lf_pad(x)];
        }
        __syncthreads();
    }
}

// ---- R1': row inverse FFT -> g_psi0; g_probe_m = g conj(T); g_T = sum_m g conj(probe_m); chain rule -----------------
template <int N> __global__ void __launch_bounds__(LF_NT) k_lf_rows_inv_mod_adj(LfParams p) {
    ADM_SMEM_DECL
    constexpr int RPC = LF_ROW_ELEMS / N, EPT = LF_ROW_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* s = (cpx*)smem_raw;
    cpx* tw = s + RPC * ld;
    lf_init_twiddles<N>(tw, tid);
    const int y0 = blockIdx.x * RPC;
    cpx tv[EPT], gt[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        tv[k] = lf_transmission(p.unknown_type, p.k1, p.sgn, p.obj[(size_t)(y0 + f) * N + x]);
        gt[k] = mk(0.f, 0.f);
    }
    for (int m = 0; m < p.M; ++m) {
        const cpx* g1 = p.in + (size_t)m * N * N;
        const cpx* pr = p.probe + (size_t)m * N * N;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
            s[f * ld + lf_pad(x)] = g1[(size_t)(y0 + f) * N + x];
        }
        __syncthreads();
        lf_inv_from<N, N>(s, ld, RPC, tw, tid);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
            const cpx g = s[f * ld + lf_pad(x)];
            const size_t gi = (size_t)(y0 + f) * N + x;
            if (p.g_probe != nullptr) p.g_probe[(size_t)m * N * N + gi] = cmulc(g, tv[k]);
            gt[k] = gt[k] + cmulc(g, pr[gi]);
        }
        __syncthreads();
    }
    if (p.g_obj != nullptr) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
            cpx o = gt[k];
            if (p.unknown_type == ADM_DELTA_BETA) {
                const cpx c = cmulc(gt[k], tv[k]);      // g_T conj(T)
                o = mk(-p.sgn * p.k1 * c.y, -p.k1 * c.x);
            }
            p.g_obj[(size_t)(y0 + f) * N + x] = o;
        }
    }
}

// ---- R2 / R2': plain row transforms of `count` frames (in place allowed) ----------------------------------------
template <int N, bool INV> __global__ void __launch_bounds__(LF_NT) k_lf_rows(const cpx* in, cpx* out) {
    ADM_SMEM_DECL
    constexpr int RPC = LF_ROW_ELEMS / N, EPT = LF_ROW_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* s = (cpx*)smem_raw;
    cpx* tw = s + RPC * ld;
    lf_init_twiddles<N>(tw, tid);
    const size_t r0 = (size_t)blockIdx.x * RPC;      // global row index over all frames
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        s[f * ld + lf_pad(x)] = in[(r0 + f) * N + x];
    }
    __syncthreads();
    if constexpr (INV) lf_inv_from<N, N>(s, ld, RPC, tw, tid);
    else lf_fwd_from<N, N>(s, ld, RPC, tw, tid);
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        out[(r0 + f) * N + x] = s[f * ld + lf_pad(x)];
    }
}

// ---- C1 / C1': column passes with the transfer functions of all distances ------------------------------------------
// grid (N / C, M).  ADJ = false: in = W1 [M][N][N], out = det [D][M][N][N];  ADJ = true: in = W2 [D][M][N][N], out = G1 [M][N][N]
template <int N, bool ADJ> __global__ void __launch_bounds__(LF_NT) k_lf_cols(LfParams p) {
    ADM_SMEM_DECL
    constexpr int C = LF_COL_ELEMS / N, EPT = LF_COL_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* a = (cpx*)smem_raw;
    cpx* bf = a + C * ld;
    cpx* tw = bf + C * ld;
    lf_init_twiddles<N>(tw, tid);
    const int x0 = blockIdx.x * C, m = blockIdx.y;
    const size_t frame = (size_t)N * N;
    if constexpr (!ADJ) {
        const cpx* w1 = p.in + (size_t)m * frame;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            a[c * ld + lf_pad(y)] = w1[(size_t)y * N + x0 + c];
        }
        __syncthreads();
        lf_fwd_from<N, N>(a, ld, C, tw, tid);
        for (int d = 0; d < p.D; ++d) {
            const cpx* h = p.H + (size_t)d * frame;
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
                bf[c * ld + lf_pad(y)] = cmul(a[c * ld + lf_pad(y)], h[(size_t)y * N + x0 + c]);
            }
            __syncthreads();
            lf_inv_from<N, N>(bf, ld, C, tw, tid);
            cpx* o = p.out + ((size_t)d * p.M + m) * frame;
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
                o[(size_t)y * N + x0 + c] = bf[c * ld + lf_pad(y)];
            }
            __syncthreads();
        }
    } else {
        cpx acc[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) acc[k] = mk(0.f, 0.f);
        for (int d = 0; d < p.D; ++d) {
            const cpx* g2 = p.in + ((size_t)d * p.M + m) * frame;
            const cpx* h = p.H + (size_t)d * frame;
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
                bf[c * ld + lf_pad(y)] = g2[(size_t)y * N + x0 + c];
            }
            __syncthreads();
            lf_fwd_from<N, N>(bf, ld, C, tw, tid);
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
                acc[k] = acc[k] + cmul(bf[c * ld + lf_pad(y)], h[(size_t)y * N + x0 + c]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            a[c * ld + lf_pad(y)] = acc[k];
        }
        __syncthreads();
        lf_inv_from<N, N>(a, ld, C, tw, tid);
        cpx* o = p.out + (size_t)m * frame;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            o[(size_t)y * N + x0 + c] = a[c * ld + lf_pad(y)];
        }
    }
}

// ---- measured-frame shift (util.py:380-397 on real data; only the real part is used, forward_model.py:1083) ----------------
// rows: real frame -> row DIF FFT (complex out)
template <int N> __global__ void __launch_bounds__(LF_NT) k_lf_rows_real_fwd(const float* in, cpx* out) {
    ADM_SMEM_DECL
    constexpr int RPC = LF_ROW_ELEMS / N, EPT = LF_ROW_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* s = (cpx*)smem_raw;
    cpx* tw = s + RPC * ld;
    lf_init_twiddles<N>(tw, tid);
    const size_t r0 = (size_t)blockIdx.x * RPC;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        s[f * ld + lf_pad(x)] = mk(in[(r0 + f) * N + x], 0.f);
    }
    __syncthreads();
    lf_fwd_from<N, N>(s, ld, RPC, tw, tid);
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        out[(r0 + f) * N + x] = s[f * ld + lf_pad(x)];
    }
}
// rows: row inverse FFT, real part out, scaled by 1/N^2
template <int N> __global__ void __launch_bounds__(LF_NT) k_lf_rows_inv_real(const cpx* in, float* out) {
    ADM_SMEM_DECL
    constexpr int RPC = LF_ROW_ELEMS / N, EPT = LF_ROW_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* s = (cpx*)smem_raw;
    cpx* tw = s + RPC * ld;
    lf_init_twiddles<N>(tw, tid);
    const size_t r0 = (size_t)blockIdx.x * RPC;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        s[f * ld + lf_pad(x)] = in[(r0 + f) * N + x];
    }
    __syncthreads();
    lf_inv_from<N, N>(s, ld, RPC, tw, tid);
    const float sc = 1.0f / ((float)N * (float)N);
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        out[(r0 + f) * N + x] = s[f * ld + lf_pad(x)].x * sc;
    }
}

// signed fftfreq (cycles / pixel) of frequency index k
ADM_DEV float lf_fftfreq(int k, int N) { return (float)(k < N / 2 ? k : k - N) / (float)N; }

// columns.  GRAD = false: column DIF FFT, multiply by exp(-2 pi i (fx sx + fy sy)), keep the shifted spectrum
// (spec, for the gradient), column inverse FFT -> out.   GRAD = true: column DIF FFT of the residual spectrum rows,
// g_shift[frame] += Re sum_k conj(G_k) spec_k (-2 pi i f(k)) / N^2 for f = fy, fx.
// `order` [N] = frequency index at each spectrum position (adm_lf_spectrum_order), grid (N / C, frames).
template <int N, bool GRAD> __global__ void __launch_bounds__(LF_NT) k_lf_cols_shift(const cpx* in, const float* shift, const int* order,
                                                                                     cpx* spec, cpx* out, float* g_shift) {
    ADM_SMEM_DECL
    constexpr int C = LF_COL_ELEMS / N, EPT = LF_COL_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* a = (cpx*)smem_raw;
    cpx* tw = a + C * ld;
    float* red = (float*)(tw + N);
    lf_init_twiddles<N>(tw, tid);
    const int x0 = blockIdx.x * C, fr = blockIdx.y;
    const size_t frame = (size_t)N * N;
    const cpx* src = in + (size_t)fr * frame;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
        a[c * ld + lf_pad(y)] = src[(size_t)y * N + x0 + c];
    }
    __syncthreads();
    lf_fwd_from<N, N>(a, ld, C, tw, tid);
    const float PI_REF = 3.14159265359f;      // adorym/constants.py (quirk 11)
    if constexpr (!GRAD) {
        const float sy = shift[2 * fr], sx = shift[2 * fr + 1];
        cpx* sp = spec + (size_t)fr * frame;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            const float fy = lf_fftfreq(order[y], N), fx = lf_fftfreq(order[x0 + c], N);
            const float ph = -2.f * PI_REF * (fx * sx + fy * sy);
            float sn, cs;
#ifndef ADM_HOST_EMU
            sincosf(ph, &sn, &cs);
#else
            sn = sinf(ph); cs = cosf(ph);
#endif
            const cpx v = cmul(a[c * ld + lf_pad(y)], mk(cs, sn));
            a[c * ld + lf_pad(y)] = v;
            sp[(size_t)y * N + x0 + c] = v;
        }
        __syncthreads();
        lf_inv_from<N, N>(a, ld, C, tw, tid);
        cpx* o = out + (size_t)fr * frame;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            o[(size_t)y * N + x0 + c] = a[c * ld + lf_pad(y)];
        }
    } else {
        const cpx* sp = spec + (size_t)fr * frame;
        float gy = 0.f, gx = 0.f;
        const float sc = -2.f * PI_REF / ((float)N * (float)N);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            const float fy = lf_fftfreq(order[y], N), fx = lf_fftfreq(order[x0 + c], N);
            const cpx g = a[c * ld + lf_pad(y)], s = sp[(size_t)y * N + x0 + c];
            // Re(conj(g) . s . (-2 pi i f)) = 2 pi f Im(conj(g) s) ; conj(g) s = (g.x s.x + g.y s.y, g.x s.y - g.y s.x)
            const float im = g.x * s.y - g.y * s.x;
            gy -= fy * im;
            gx -= fx * im;
        }
        // block reduction, one atomic pair per CTA
        red[tid] = gy * sc;
        red[LF_NT + tid] = gx * sc;
        __syncthreads();
        for (int st = LF_NT / 2; st > 0; st >>= 1) {
            if (tid < st) { red[tid] += red[tid + st]; red[LF_NT + tid] += red[LF_NT + tid + st]; }
            __syncthreads();
        }
        if (tid == 0) {
            atomicAdd(&g_shift[2 * fr], red[0]);
            atomicAdd(&g_shift[2 * fr + 1], red[LF_NT]);
        }
    }
}

// ---- ptychography with probes the shared-memory FFT cannot hold (128 .. 2048 pixels): batched patch steps -------------------
// One slice step of propagate.py:208-258 for B*M wavefields at once, every pattern b reading ITS patch of the slice-major
// transmission volume T' (origin pos[b], vacuum outside the object: util.py:1327-1351, forward_model.py:313-331):
//   psi <- IFFT2(H . FFT2(psi (1 + T'_patch)))      (H == NULL: modulation only -- the last slice, propagate.py:253)
// as   P1 [modulation + row FFT]  ->  k_lf_cols (column FFT . H . column IFFT)  ->  k_lf_rows (row IFFT),
// and the adjoint in the reverse order; dL/dT = g conj(psi_in) is scatter-added into the gradient volume (red.add, the
// patches of a scan overlap).
struct LfPatch {
    const cpx* tvol;     // [Nz][Y][Xp]
    const int* pos;      // [B][2] (y, x)
    const cpx* psi_in;   // [B*M][N][N]
    const cpx* g_in;     // adjoint: incoming gradient (row-transformed when FFT)
    cpx* out;            // forward: [B*M][N][N]
    cpx* g_psi;          // adjoint: dL/d psi_in [B*M][N][N]
    cpx* g_tvol;         // adjoint: [Nz][Y + gm][gP], accumulated
    int Y, X, Xp, z, M, gm, gP;
};

ADM_DEV cpx lf_patch_tp(const LfPatch& p, int gy, int gx) {
    const bool inb = (unsigned)gy < (unsigned)p.Y && (unsigned)gx < (unsigned)p.X;
    return inb ? p.tvol[((size_t)p.z * p.Y + gy) * p.Xp + gx] : mk(0.f, 0.f);
}

template <int N, bool FFT> __global__ void __launch_bounds__(LF_NT) k_lf_patch_mod(LfPatch p) {
    ADM_SMEM_DECL
    constexpr int RPC = LF_ROW_ELEMS / N, EPT = LF_ROW_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* s = (cpx*)smem_raw;
    cpx* tw = s + RPC * ld;
    if (FFT) lf_init_twiddles<N>(tw, tid);
    const int y0 = blockIdx.x * RPC, fr = blockIdx.y, b = fr / p.M;
    const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
    const cpx* in = p.psi_in + (size_t)fr * N * N;
    cpx* o = p.out + (size_t)fr * N * N;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        const cpx v = modulate(in[(size_t)(y0 + f) * N + x], lf_patch_tp(p, py + y0 + f, px + x));
        if (FFT) s[f * ld + lf_pad(x)] = v;
        else o[(size_t)(y0 + f) * N + x] = v;
    }
    if (FFT) {
        __syncthreads();
        lf_fwd_from<N, N>(s, ld, RPC, tw, tid);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
            o[(size_t)(y0 + f) * N + x] = s[f * ld + lf_pad(x)];
        }
    }
}

template <int N, bool FFT> __global__ void __launch_bounds__(LF_NT) k_lf_patch_mod_adj(LfPatch p) {
    ADM_SMEM_DECL
    constexpr int RPC = LF_ROW_ELEMS / N, EPT = LF_ROW_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* s = (cpx*)smem_raw;
    cpx* tw = s + RPC * ld;
    if (FFT) lf_init_twiddles<N>(tw, tid);
    const int y0 = blockIdx.x * RPC, fr = blockIdx.y, b = fr / p.M;
    const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
    const cpx* gi = p.g_in + (size_t)fr * N * N;
    const cpx* in = p.psi_in + (size_t)fr * N * N;
    if (FFT) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
            s[f * ld + lf_pad(x)] = gi[(size_t)(y0 + f) * N + x];
        }
        __syncthreads();
        lf_inv_from<N, N>(s, ld, RPC, tw, tid);
    }
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int idx = tid + k * LF_NT, f = idx / N, x = idx % N;
        const size_t e = (size_t)(y0 + f) * N + x;
        const cpx g = FFT ? s[f * ld + lf_pad(x)] : gi[e];
        const int gy = py + y0 + f, gx = px + x;
        if (p.g_psi != nullptr) p.g_psi[(size_t)fr * N * N + e] = modulate_conj(g, lf_patch_tp(p, gy, gx));
        if (p.g_tvol != nullptr && (unsigned)gy < (unsigned)p.Y && (unsigned)gx < (unsigned)p.X) {
            const cpx c = cmulc(g, in[e]);                     // dL/dT = g conj(psi_in)
            lf_red2(p.g_tvol + ((size_t)p.z * (p.Y + p.gm) + gy + p.gm) * p.gP + gx + p.gm, c.x, c.y);
        }
    }
}

// far field (propagate.py:263-270, sign_convention = +1): det = fftshift(FFT2(psi)) . scale, natural order.
//   ADJ = false: in = row-transformed frames (k_lf_rows<N, false>), column DIF FFT, scatter to the shifted natural position
//   ADJ = true : in = dL/d det (natural, shifted), gather into the permuted order, column inverse (= conjugate transpose)
// grid (N / C, frames); order [N] = frequency index at each spectrum position.
template <int N, bool ADJ> __global__ void __launch_bounds__(LF_NT) k_lf_cols_far(const cpx* in, const int* order, float scale,
                                                                                   cpx* out) {
    ADM_SMEM_DECL
    constexpr int C = LF_COL_ELEMS / N, EPT = LF_COL_ELEMS / LF_NT;
    const int ld = lf_ld(N), tid = threadIdx.x;
    cpx* a = (cpx*)smem_raw;
    cpx* tw = a + C * ld;
    lf_init_twiddles<N>(tw, tid);
    const int x0 = blockIdx.x * C, fr = blockIdx.y;
    const size_t frame = (size_t)N * N;
    const cpx* src = in + (size_t)fr * frame;
    cpx* dst = out + (size_t)fr * frame;
    if constexpr (!ADJ) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            a[c * ld + lf_pad(y)] = src[(size_t)y * N + x0 + c];
        }
        __syncthreads();
        lf_fwd_from<N, N>(a, ld, C, tw, tid);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            const int fy = (order[y] + N / 2) % N, fx = (order[x0 + c] + N / 2) % N;
            dst[(size_t)fy * N + fx] = cscale(a[c * ld + lf_pad(y)], scale);
        }
    } else {
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            const int fy = (order[y] + N / 2) % N, fx = (order[x0 + c] + N / 2) % N;
            a[c * ld + lf_pad(y)] = cscale(src[(size_t)fy * N + fx], scale);
        }
        __syncthreads();
        lf_inv_from<N, N>(a, ld, C, tw, tid);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int idx = tid + k * LF_NT, c = idx % C, y = idx / C;
            dst[(size_t)y * N + x0 + c] = a[c * ld + lf_pad(y)];
        }
    }
}

// ---- detector magnitude + mismatch + residuals for large frames: grid (chunks, frames), deterministic 2-stage sum ------
#define LF_LOSS_CHUNKS 64
__global__ void __launch_bounds__(LF_NT) k_lf_detector_loss(int M, int npix, int loss_type, int raw_type, float pm, float inv_n,
                                                            const cpx* det, const float* meas, float* pred, cpx* g_det,
                                                            float* g_meas, float* partial) {
#ifndef ADM_HOST_EMU
    __shared__ float red[LF_NT];
#else
    float* red = (float*)emu_smem;
#endif
    const int b = blockIdx.y;
    const cpx* d = det + (size_t)b * M * npix;
    float acc = 0.f;
    for (int i = blockIdx.x * LF_NT + threadIdx.x; i < npix; i += LF_NT * LF_LOSS_CHUNKS) {
        float s = 0.f;
        for (int m = 0; m < M; ++m) {
            const cpx v = d[(size_t)m * npix + i];
            s += v.x * v.x + v.y * v.y;
        }
        const float a = sqrtf(s);
        pred[(size_t)b * npix + i] = a;
        const float y = meas[(size_t)b * npix + i];
        float dlda, dldy;
        if (loss_type == ADM_LOSS_LSQ) {
            const float ay = fabsf(y);
            const float mm = (raw_type == ADM_RAW_MAGNITUDE) ? ay : sqrtf(ay);
            const float r = a - mm;
            acc += r * r;
            dlda = 2.f * r * inv_n;
            const float sg = (y > 0.f) ? 1.f : ((y < 0.f) ? -1.f : 0.f);
            dldy = (raw_type == ADM_RAW_MAGNITUDE) ? -dlda * sg : ((ay > 0.f) ? -dlda * sg * 0.5f / mm : 0.f);
        } else {
            const float ay = fabsf(y);
            const float Im = (raw_type == ADM_RAW_MAGNITUDE) ? ay * ay : ay;
            const float a2 = a * a * pm;
            const float lg = logf(a2);
            acc += a2 - Im * pm * lg;
            dlda = (a > 0.f) ? (2.f * pm * a - 2.f * pm * Im / a) * inv_n : 0.f;
            const float sg = (y > 0.f) ? 1.f : ((y < 0.f) ? -1.f : 0.f);
            dldy = -pm * lg * inv_n * ((raw_type == ADM_RAW_MAGNITUDE) ? 2.f * ay * sg : sg);
        }
        if (g_meas != nullptr) g_meas[(size_t)b * npix + i] = dldy;
        if (g_det != nullptr) {
            const float c = (a > 0.f) ? dlda / a : 0.f;
            for (int m = 0; m < M; ++m) {
                const cpx v = d[(size_t)m * npix + i];
                g_det[((size_t)b * M + m) * npix + i] = mk(c * v.x, c * v.y);
            }
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = LF_NT / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[b * LF_LOSS_CHUNKS + blockIdx.x] = red[0];
}
__global__ void k_lf_loss_reduce(const float* partial, float* loss_per_pattern) {
    // one warp-sized CTA per frame, fixed summation order
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < LF_LOSS_CHUNKS; ++i) s += partial[blockIdx.x * LF_LOSS_CHUNKS + i];
        loss_per_pattern[blockIdx.x] = s;
    }
}

// ================================================================================================
// host side
// ================================================================================================
template <int N> struct LfSizes {
    static constexpr int RPC = LF_ROW_ELEMS / N, C = LF_COL_ELEMS / N;
    static size_t rows() { return sizeof(cpx) * ((size_t)RPC * lf_ld(N) + N) + 16; }
    static size_t cols2() { return sizeof(cpx) * ((size_t)2 * C * lf_ld(N) + N) + 16; }
    static size_t cols1() { return sizeof(cpx) * ((size_t)C * lf_ld(N) + N) + 2 * LF_NT * sizeof(float) + 16; }
};

static bool lf_supported(int N) { return N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048; }

#define LF_DISPATCH(N, ...)                                        \
    switch (N) {                                                   \
        case 128: { constexpr int LN = 128; __VA_ARGS__ } break;   \
        case 256: { constexpr int LN = 256; __VA_ARGS__ } break;   \
        case 512: { constexpr int LN = 512; __VA_ARGS__ } break;   \
        case 1024: { constexpr int LN = 1024; __VA_ARGS__ } break; \
        case 2048: { constexpr int LN = 2048; __VA_ARGS__ } break; \
        default: return ADM_ERR_UNSUPPORTED;                       \
    }

extern "C" int adm_lf_supported(int N) { return lf_supported(N) ? 1 : 0; }

extern "C" int adm_lf_spectrum_order(int N, int* freq_of_pos) {
    if (!freq_of_pos) return ADM_ERR_ARG;
    if (!lf_supported(N)) return ADM_ERR_UNSUPPORTED;
    for (int p = 0; p < N; ++p) freq_of_pos[p] = lf_freq_of_pos(N, p);
    return ADM_OK;
}

static int lf_fill(const adm_multidist_cfg* c, LfParams& p) {
    if (!c) return ADM_ERR_ARG;
    if (!lf_supported(c->N)) return ADM_ERR_UNSUPPORTED;
    if (c->M < 1 || c->D < 1) return ADM_ERR_ARG;
    if (c->sign_convention != 1 && c->sign_convention != -1) return ADM_ERR_ARG;
    if (c->unknown_type != ADM_DELTA_BETA && c->unknown_type != ADM_REAL_IMAG) return ADM_ERR_ARG;
    p.M = c->M; p.D = c->D; p.unknown_type = c->unknown_type; p.k1 = c->k1; p.sgn = (float)c->sign_convention;
    p.obj = nullptr; p.probe = nullptr; p.H = nullptr; p.in = nullptr; p.out = nullptr; p.g_obj = nullptr; p.g_probe = nullptr;
    return ADM_OK;
}

extern "C" int adm_multidist_fwd(const adm_multidist_cfg* cfg, const float* obj, const float* probe, const float* Hprep,
                                 float* work1, float* det, void* stream) {
    LfParams p;
    int rc = lf_fill(cfg, p);
    if (rc) return rc;
    if (!obj || !probe || !Hprep || !work1 || !det) return ADM_ERR_ARG;
    const int N = cfg->N, M = cfg->M, D = cfg->D;
    LF_DISPATCH(N, {
        if (set_smem(k_lf_rows_mod_fwd<LN>, LfSizes<LN>::rows()) || set_smem(k_lf_cols<LN, false>, LfSizes<LN>::cols2()) ||
            set_smem(k_lf_rows<LN, true>, LfSizes<LN>::rows()))
            return ADM_ERR_CUDA;
        LfParams a = p;
        a.obj = (const cpx*)obj; a.probe = (const cpx*)probe; a.out = (cpx*)work1;
        ADM_LAUNCH(k_lf_rows_mod_fwd<LN>, LN / LfSizes<LN>::RPC, LF_NT, LfSizes<LN>::rows(), stream, a);
        LfParams b = p;
        b.in = (const cpx*)work1; b.H = (const cpx*)Hprep; b.out = (cpx*)det;
        ADM_LAUNCH((k_lf_cols<LN, false>), dim3(LN / LfSizes<LN>::C, M), LF_NT, LfSizes<LN>::cols2(), stream, b);
        ADM_LAUNCH((k_lf_rows<LN, true>), D * M * (LN / LfSizes<LN>::RPC), LF_NT, LfSizes<LN>::rows(), stream,
                   (const cpx*)det, (cpx*)det);
    })
    return ADM_OK;
}

extern "C" int adm_multidist_bwd(const adm_multidist_cfg* cfg, const float* obj, const float* probe, const float* HprepC,
                                 const float* g_det, float* work1, float* work2, float* g_obj, float* g_probe, void* stream) {
    LfParams p;
    int rc = lf_fill(cfg, p);
    if (rc) return rc;
    if (!obj || !probe || !HprepC || !g_det || !work1 || !work2) return ADM_ERR_ARG;
    const int N = cfg->N, M = cfg->M, D = cfg->D;
    LF_DISPATCH(N, {
        if (set_smem(k_lf_rows_inv_mod_adj<LN>, LfSizes<LN>::rows()) || set_smem(k_lf_cols<LN, true>, LfSizes<LN>::cols2()) ||
            set_smem(k_lf_rows<LN, false>, LfSizes<LN>::rows()))
            return ADM_ERR_CUDA;
        ADM_LAUNCH((k_lf_rows<LN, false>), D * M * (LN / LfSizes<LN>::RPC), LF_NT, LfSizes<LN>::rows(), stream,
                   (const cpx*)g_det, (cpx*)work2);
        LfParams b = p;
        b.in = (const cpx*)work2; b.H = (const cpx*)HprepC; b.out = (cpx*)work1;
        ADM_LAUNCH((k_lf_cols<LN, true>), dim3(LN / LfSizes<LN>::C, M), LF_NT, LfSizes<LN>::cols2(), stream, b);
        LfParams a = p;
        a.obj = (const cpx*)obj; a.probe = (const cpx*)probe; a.in = (const cpx*)work1;
        a.g_obj = (cpx*)g_obj; a.g_probe = (cpx*)g_probe;
        ADM_LAUNCH(k_lf_rows_inv_mod_adj<LN>, LN / LfSizes<LN>::RPC, LF_NT, LfSizes<LN>::rows(), stream, a);
    })
    return ADM_OK;
}

static int lf_patch_fill(LfPatch& q, int B, int M, int Y, int X, int z, const float* tvol, const int* pos) {
    if (B < 1 || M < 1 || Y < 1 || X < 1 || z < 0 || !tvol || !pos) return ADM_ERR_ARG;
    if ((long long)B * M > 65535) return ADM_ERR_UNSUPPORTED;
    q.tvol = (const cpx*)tvol; q.pos = pos; q.psi_in = nullptr; q.g_in = nullptr; q.out = nullptr; q.g_psi = nullptr;
    q.g_tvol = nullptr; q.Y = Y; q.X = X; q.Xp = ADM_TVOL_PITCH(X); q.z = z; q.M = M; q.gm = 0; q.gP = 0;
    return ADM_OK;
}

extern "C" int adm_lf_patch_step_fwd(int N, int B, int M, int Y, int X, int z, const float* tvol, const int* pos,
                                     const float* psi_in, const float* Hprep, float* work, float* psi_out, void* stream) {
    LfPatch q;
    int rc = lf_patch_fill(q, B, M, Y, X, z, tvol, pos);
    if (rc) return rc;
    if (!psi_in || !psi_out || (Hprep && !work)) return ADM_ERR_ARG;
    const int F = B * M;
    q.psi_in = (const cpx*)psi_in;
    LF_DISPATCH(N, {
        if (!Hprep) {
            q.out = (cpx*)psi_out;
            ADM_LAUNCH((k_lf_patch_mod<LN, false>), dim3(LN / LfSizes<LN>::RPC, F), LF_NT, 16, stream, q);
        } else {
            if (set_smem(k_lf_patch_mod<LN, true>, LfSizes<LN>::rows()) || set_smem(k_lf_cols<LN, false>, LfSizes<LN>::cols2()) ||
                set_smem(k_lf_rows<LN, true>, LfSizes<LN>::rows()))
                return ADM_ERR_CUDA;
            q.out = (cpx*)work;
            ADM_LAUNCH((k_lf_patch_mod<LN, true>), dim3(LN / LfSizes<LN>::RPC, F), LF_NT, LfSizes<LN>::rows(), stream, q);
            LfParams b;
            b.obj = nullptr; b.probe = nullptr; b.g_obj = nullptr; b.g_probe = nullptr; b.unknown_type = 0; b.k1 = 0.f; b.sgn = 1.f;
            b.M = F; b.D = 1; b.in = (const cpx*)work; b.H = (const cpx*)Hprep; b.out = (cpx*)psi_out;
            ADM_LAUNCH((k_lf_cols<LN, false>), dim3(LN / LfSizes<LN>::C, F), LF_NT, LfSizes<LN>::cols2(), stream, b);
            ADM_LAUNCH((k_lf_rows<LN, true>), F * (LN / LfSizes<LN>::RPC), LF_NT, LfSizes<LN>::rows(), stream,
                       (const cpx*)psi_out, (cpx*)psi_out);
        }
    })
    return ADM_OK;
}

extern "C" int adm_lf_patch_step_bwd(int N, int B, int M, int Y, int X, int z, const float* tvol, const int* pos,
                                     const float* psi_in, const float* HprepC, const float* g_out, float* work1, float* work2,
                                     float* g_psi, float* g_tvol, int g_margin, void* stream) {
    LfPatch q;
    int rc = lf_patch_fill(q, B, M, Y, X, z, tvol, pos);
    if (rc) return rc;
    if (!psi_in || !g_out || g_margin < 0 || (HprepC && (!work1 || !work2))) return ADM_ERR_ARG;
    const int F = B * M;
    q.psi_in = (const cpx*)psi_in; q.g_psi = (cpx*)g_psi; q.g_tvol = (cpx*)g_tvol;
    q.gm = g_margin; q.gP = ADM_GVOL_PITCH(X, g_margin);
    LF_DISPATCH(N, {
        if (!HprepC) {
            q.g_in = (const cpx*)g_out;
            ADM_LAUNCH((k_lf_patch_mod_adj<LN, false>), dim3(LN / LfSizes<LN>::RPC, F), LF_NT, 16, stream, q);
        } else {
            if (set_smem(k_lf_patch_mod_adj<LN, true>, LfSizes<LN>::rows()) || set_smem(k_lf_cols<LN, true>, LfSizes<LN>::cols2()) ||
                set_smem(k_lf_rows<LN, false>, LfSizes<LN>::rows()))
                return ADM_ERR_CUDA;
            ADM_LAUNCH((k_lf_rows<LN, false>), F * (LN / LfSizes<LN>::RPC), LF_NT, LfSizes<LN>::rows(), stream,
                       (const cpx*)g_out, (cpx*)work2);
            LfParams b;
            b.obj = nullptr; b.probe = nullptr; b.g_obj = nullptr; b.g_probe = nullptr; b.unknown_type = 0; b.k1 = 0.f; b.sgn = 1.f;
            b.M = F; b.D = 1; b.in = (const cpx*)work2; b.H = (const cpx*)HprepC; b.out = (cpx*)work1;
            ADM_LAUNCH((k_lf_cols<LN, true>), dim3(LN / LfSizes<LN>::C, F), LF_NT, LfSizes<LN>::cols2(), stream, b);
            q.g_in = (const cpx*)work1;
            ADM_LAUNCH((k_lf_patch_mod_adj<LN, true>), dim3(LN / LfSizes<LN>::RPC, F), LF_NT, LfSizes<LN>::rows(), stream, q);
        }
    })
    return ADM_OK;
}

extern "C" int adm_lf_farfield_fwd(int N, int frames, const float* in, const int* order, float scale, float* work, float* out,
                                   void* stream) {
    if (!in || !order || !work || !out || frames < 1 || frames > 65535) return ADM_ERR_ARG;
    LF_DISPATCH(N, {
        if (set_smem(k_lf_rows<LN, false>, LfSizes<LN>::rows()) || set_smem(k_lf_cols_far<LN, false>, LfSizes<LN>::cols1()))
            return ADM_ERR_CUDA;
        ADM_LAUNCH((k_lf_rows<LN, false>), frames * (LN / LfSizes<LN>::RPC), LF_NT, LfSizes<LN>::rows(), stream, (const cpx*)in,
                   (cpx*)work);
        ADM_LAUNCH((k_lf_cols_far<LN, false>), dim3(LN / LfSizes<LN>::C, frames), LF_NT, LfSizes<LN>::cols1(), stream,
                   (const cpx*)work, order, scale, (cpx*)out);
    })
    return ADM_OK;
}

extern "C" int adm_lf_farfield_bwd(int N, int frames, const float* g_out, const int* order, float scale, float* work,
                                   float* g_in, void* stream) {
    if (!g_out || !order || !work || !g_in || frames < 1 || frames > 65535) return ADM_ERR_ARG;
    LF_DISPATCH(N, {
        if (set_smem(k_lf_rows<LN, true>, LfSizes<LN>::rows()) || set_smem(k_lf_cols_far<LN, true>, LfSizes<LN>::cols1()))
            return ADM_ERR_CUDA;
        ADM_LAUNCH((k_lf_cols_far<LN, true>), dim3(LN / LfSizes<LN>::C, frames), LF_NT, LfSizes<LN>::cols1(), stream,
                   (const cpx*)g_out, order, scale, (cpx*)work);
        ADM_LAUNCH((k_lf_rows<LN, true>), frames * (LN / LfSizes<LN>::RPC), LF_NT, LfSizes<LN>::rows(), stream, (const cpx*)work,
                   (cpx*)g_in);
    })
    return ADM_OK;
}

extern "C" int adm_frame_shift_fwd(int N, int frames, const float* data, const float* shift, const int* order, float* work,
                                   float* spec, float* out, void* stream) {
    if (!data || !shift || !order || !work || !spec || !out || frames < 1) return ADM_ERR_ARG;
    LF_DISPATCH(N, {
        if (set_smem(k_lf_rows_real_fwd<LN>, LfSizes<LN>::rows()) || set_smem(k_lf_cols_shift<LN, false>, LfSizes<LN>::cols1()) ||
            set_smem(k_lf_rows_inv_real<LN>, LfSizes<LN>::rows()))
            return ADM_ERR_CUDA;
        const int rg = frames * (LN / LfSizes<LN>::RPC);
        ADM_LAUNCH(k_lf_rows_real_fwd<LN>, rg, LF_NT, LfSizes<LN>::rows(), stream, data, (cpx*)work);
        ADM_LAUNCH((k_lf_cols_shift<LN, false>), dim3(LN / LfSizes<LN>::C, frames), LF_NT, LfSizes<LN>::cols1(), stream,
                   (const cpx*)work, shift, order, (cpx*)spec, (cpx*)work, (float*)nullptr);
        ADM_LAUNCH(k_lf_rows_inv_real<LN>, rg, LF_NT, LfSizes<LN>::rows(), stream, (const cpx*)work, out);
    })
    return ADM_OK;
}

extern "C" int adm_frame_shift_bwd(int N, int frames, const float* g_out, const float* spec, const int* order, float* work,
                                   float* g_shift, void* stream) {
    if (!g_out || !spec || !order || !work || !g_shift || frames < 1) return ADM_ERR_ARG;
    LF_DISPATCH(N, {
        if (set_smem(k_lf_rows_real_fwd<LN>, LfSizes<LN>::rows()) || set_smem(k_lf_cols_shift<LN, true>, LfSizes<LN>::cols1()))
            return ADM_ERR_CUDA;
        const int rg = frames * (LN / LfSizes<LN>::RPC);
        ADM_LAUNCH(k_lf_rows_real_fwd<LN>, rg, LF_NT, LfSizes<LN>::rows(), stream, g_out, (cpx*)work);
        ADM_LAUNCH((k_lf_cols_shift<LN, true>), dim3(LN / LfSizes<LN>::C, frames), LF_NT, LfSizes<LN>::cols1(), stream,
                   (const cpx*)work, (const float*)nullptr, order, (cpx*)spec, (cpx*)nullptr, g_shift);
    })
    return ADM_OK;
}

// complex frames: the same column kernels behind plain complex row transforms; the inverse is scaled by 1 / N^2 in place
__global__ void k_lf_scale(cpx* a, float sc, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = cscale(a[i], sc);
}

extern "C" int adm_frame_shift_complex_fwd(int N, int frames, const float* in, const float* shift, const int* order,
                                           float* work, float* spec, float* out, void* stream) {
    if (!in || !shift || !order || !work || !spec || !out || frames < 1 || frames > 65535) return ADM_ERR_ARG;
    LF_DISPATCH(N, {
        if (set_smem(k_lf_rows<LN, false>, LfSizes<LN>::rows()) || set_smem(k_lf_cols_shift<LN, false>, LfSizes<LN>::cols1()) ||
            set_smem(k_lf_rows<LN, true>, LfSizes<LN>::rows()))
            return ADM_ERR_CUDA;
        const int rg = frames * (LN / LfSizes<LN>::RPC);
        ADM_LAUNCH((k_lf_rows<LN, false>), rg, LF_NT, LfSizes<LN>::rows(), stream, (const cpx*)in, (cpx*)work);
        ADM_LAUNCH((k_lf_cols_shift<LN, false>), dim3(LN / LfSizes<LN>::C, frames), LF_NT, LfSizes<LN>::cols1(), stream,
                   (const cpx*)work, shift, order, (cpx*)spec, (cpx*)work, (float*)nullptr);
        ADM_LAUNCH((k_lf_rows<LN, true>), rg, LF_NT, LfSizes<LN>::rows(), stream, (const cpx*)work, (cpx*)out);
        const long long n = (long long)frames * LN * LN;
        ADM_LAUNCH(k_lf_scale, (unsigned)((n + 255) / 256), 256, 0, stream, (cpx*)out, 1.0f / ((float)LN * (float)LN), n);
    })
    return ADM_OK;
}

extern "C" int adm_frame_shift_complex_bwd(int N, int frames, const float* g_out, const float* spec, const int* order,
                                           float* work, float* g_shift, void* stream) {
    if (!g_out || !spec || !order || !work || !g_shift || frames < 1 || frames > 65535) return ADM_ERR_ARG;
    LF_DISPATCH(N, {
        if (set_smem(k_lf_rows<LN, false>, LfSizes<LN>::rows()) || set_smem(k_lf_cols_shift<LN, true>, LfSizes<LN>::cols1()))
            return ADM_ERR_CUDA;
        const int rg = frames * (LN / LfSizes<LN>::RPC);
        ADM_LAUNCH((k_lf_rows<LN, false>), rg, LF_NT, LfSizes<LN>::rows(), stream, (const cpx*)g_out, (cpx*)work);
        ADM_LAUNCH((k_lf_cols_shift<LN, true>), dim3(LN / LfSizes<LN>::C, frames), LF_NT, LfSizes<LN>::cols1(), stream,
                   (const cpx*)work, (const float*)nullptr, order, (cpx*)spec, (cpx*)nullptr, g_shift);
    })
    return ADM_OK;
}

extern "C" int adm_detector_loss_large(int B, int M, long long npix, int loss_type, int raw_data_type, float poisson_multiplier,
                                       float inv_n, const float* det, const float* meas, float* pred, float* g_det,
                                       float* g_meas, float* partial, float* loss_per_pattern, void* stream) {
    if (B < 1 || M < 1 || npix < 1 || !det || !meas || !pred || !partial || !loss_per_pattern) return ADM_ERR_ARG;
    if (loss_type != ADM_LOSS_LSQ && loss_type != ADM_LOSS_POISSON) return ADM_ERR_ARG;
    if (raw_data_type != ADM_RAW_MAGNITUDE && raw_data_type != ADM_RAW_INTENSITY) return ADM_ERR_ARG;
    ADM_LAUNCH(k_lf_detector_loss, dim3(LF_LOSS_CHUNKS, B), LF_NT, LF_NT * sizeof(float), stream, M, (int)npix, loss_type,
               raw_data_type, poisson_multiplier, inv_n, (const cpx*)det, meas, pred, (cpx*)g_det, g_meas, partial);
    ADM_LAUNCH(k_lf_loss_reduce, B, 32, 0, stream, (const float*)partial, loss_per_pattern);
    return ADM_OK;
}
