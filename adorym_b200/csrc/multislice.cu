// Fused multislice forward / adjoint kernels (sm_100a).
//
// One CTA owns one wavefield (pattern b, probe mode m).  The wavefield psi[P][P] lives in shared memory for the
// whole traversal of the Nz slices; per slice the only HBM traffic is the (delta,beta) patch tile (prefetched
// with cp.async, zero-filled outside the object = vacuum padding) and the psi checkpoint the adjoint consumes.
//
// 2-D FFT of size P = 8*R2 (R2 = 8 or 9) as an in-place mixed-radix Cooley-Tukey over shared memory:
//   1-D forward:  pass A  radix-8 over n1 (x[R2 n1 + n2], stride R2) + twiddle W_P^{n2 k1}, stored at R2 k1 + n2
//                 pass B  radix-R2 over n2 (contiguous R2 k1 + n2)               -> X[k1 + 8 k2] at R2 k1 + k2
//   1-D inverse:  pass B^-1 then pass A^-1 on the same index sets, conjugate twiddles, unnormalised.
//   The spectrum therefore sits in a permuted order, perm(p) = p / R2 + 8 (p % R2); the transfer function is
//   pre-permuted (adm_prepare_transfer_function), so no reordering pass exists.
// Because pass B, the H multiply and pass B^-1 act on the same R2 points, and pass A^-1, the slice modulation
// and the next slice's pass A act on the same 8 points, a slice costs SIX shared-memory round trips:
//   P1 [Ay^-1 . t(delta,beta) . Ay]   P2 By   P3 Ax   P4 [Bx . H . Bx^-1]   P5 Ax^-1   P6 By^-1
// y-passes run with lanes along x (row stride S = P+1 complex keeps x-passes, lanes along y, conflict free).
//
// Reference behaviour reproduced: adorym/propagate.py:208-258 (slice loop, last slice not propagated :253),
// :241 transmission, :255 convolution, :263-280 detector transform; forward_model.py:313-331 + util.py:1327-1351
// gather with vacuum padding.  The adjoint replaces the autograd tape (wrappers.py:300-331).
#include "fft_passes.cuh"
#include "../../include/adorym_b200.h"

struct MsParams {
    const cpx* rot_obj;
    const cpx* probe;
    const int* pos;
    const cpx* Hp;      // prepared in-object transfer function (conj version for the adjoint)
    const cpx* Hf;      // prepared free-space transfer function (near field) or null
    cpx* ckpt;
    cpx* det;           // fwd: out ; bwd: g_det in
    cpx* g_obj;
    cpx* g_probe;
    int Nz, Y, X, B, M;
    float k1, sgn;
    int unknown_type, free_prop, probe_per_pattern;
    float ff_scale;
    int ff_conj;
};

// ---- cp.async helpers -------------------------------------------------------------------------
ADM_DEV void async_copy8(cpx* smem_dst, const cpx* gsrc, bool valid) {
#ifndef ADM_HOST_EMU
    unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    int nbytes = valid ? 8 : 0;      // src-size 0 -> the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(saddr), "l"(gsrc), "r"(nbytes) : "memory");
#else
    *smem_dst = valid ? *gsrc : mk(0.f, 0.f);
#endif
}
ADM_DEV void async_commit() {
#ifndef ADM_HOST_EMU
    asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
ADM_DEV void async_wait_all() {
#ifndef ADM_HOST_EMU
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}
ADM_DEV void red_add2(cpx* addr, float a, float b) {
#ifndef ADM_HOST_EMU
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a), "f"(b) : "memory");
#else
    atomicAdd(&addr->x, a);
    atomicAdd(&addr->y, b);
#endif
}

// transmission function of one voxel (propagate.py:241 / :246); oob real_imag voxels are vacuum (1, 0)
ADM_DEV cpx transmission(cpx o, bool inb, int unknown_type, float k1, float sgn) {
    if (unknown_type == ADM_DELTA_BETA) {
        float e = expf(-k1 * o.y);
        float s, c;
        adm_sincos(-sgn * k1 * o.x, &s, &c);
        return mk(e * c, e * s);
    }
    return inb ? o : mk(1.f, 0.f);
}

// ================================================================================================
// Forward
// ================================================================================================
template <class C> __global__ void __launch_bounds__(C::NTHREADS) k_multislice_fwd(MsParams p) {
    ADM_SMEM_DECL
    constexpr int P = C::P, R2 = C::R2, S = C::S;
    cpx* psi = (cpx*)smem_raw;
    cpx* tile = psi + P * S;
    cpx* tw = tile + P * P;
    const int t = threadIdx.x;
    const bool active = t < C::NACT;
    const int q = t / P, l = t % P;
    const int wf = blockIdx.x;              // wavefield index b * M + m
    const int b = wf / p.M, m = wf % p.M;
    const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
    const int gx = px + l;
    const bool x_in = (gx >= 0) && (gx < p.X);
    const cpx* probe = p.probe + (size_t)(p.probe_per_pattern ? wf : m) * P * P;
    cpx* ckpt = p.ckpt + (size_t)wf * p.Nz * P * P;
    cpx* det = p.det + (size_t)wf * P * P;
    const size_t slice_stride = (size_t)p.Y * p.X;

    init_twiddles<C>(tw, t);
    // prefetch the (delta,beta) tile of slice 0: every thread copies exactly the elements it will consume
    if (active) {
        for (int g = q; g < R2; g += C::Q)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int y = R2 * j + g, gy = py + y;
                bool inb = x_in && gy >= 0 && gy < p.Y;
                const cpx* src = inb ? p.rot_obj + (size_t)gy * p.X + gx : p.rot_obj;
                async_copy8(&tile[y * P + l], src, inb);
            }
        async_commit();
    }
    __syncthreads();

    for (int z = 0; z < p.Nz; ++z) {
        const bool last = (z == p.Nz - 1);
        const bool go_on = !last || p.free_prop != ADM_PROP_NONE;
        // ---- P1: [Ay^-1] . checkpoint . transmission . [Ay] --------------------------------------------
        if (active) {
            async_wait_all();
            for (int g = q; g < R2; g += C::Q) {
                cpx v[8];
                if (z == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = probe[(R2 * j + g) * P + l];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = psi[(R2 * j + g) * S + l];
                    dft8<true>(v);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    int y = R2 * j + g, gy = py + y;
                    bool inb = x_in && gy >= 0 && gy < p.Y;
                    ckpt[(size_t)z * P * P + y * P + l] = v[j];
                    cpx o = tile[y * P + l];
                    v[j] = cmul(v[j], transmission(o, inb, p.unknown_type, p.k1, p.sgn));
                    if (!last) {
                        const cpx* src = inb ? p.rot_obj + (size_t)(z + 1) * slice_stride + (size_t)gy * p.X + gx : p.rot_obj;
                        async_copy8(&tile[y * P + l], src, inb);
                    }
                }
                if (go_on) {
                    if (last && p.ff_conj) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = cconj(v[j]);
                    }
                    dft8<false>(v);
#pragma unroll
                    for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw[g * k]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) psi[(R2 * k + g) * S + l] = v[k];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) det[(R2 * j + g) * P + l] = v[j];   // exit wave, natural order
                }
            }
            async_commit();
        }
        __syncthreads();
        if (!go_on) break;
        if (last && p.free_prop == ADM_PROP_FAR) {
            // far field: finish the forward 2-D FFT (By, Ax, Bx) and store fftshift-ed, natural order
            if (active) pass_B_fwd<C, false>(psi, q, l);
            __syncthreads();
            if (active) pass_A_fwd<C, true>(psi, tw, q, l);
            __syncthreads();
            if (active) {
                const int fy = (perm_freq<C>(l) + P / 2) % P;
                for (int g = q; g < C::R1; g += C::Q) {
                    cpx v[R2];
#pragma unroll
                    for (int n = 0; n < R2; ++n) v[n] = psi[l * S + R2 * g + n];
                    DftR<R2, false>::run(v);
#pragma unroll
                    for (int k = 0; k < R2; ++k) {
                        int fx = (g + C::R1 * k + P / 2) % P;
                        cpx o = cscale(v[k], p.ff_scale);
                        det[fy * P + fx] = p.ff_conj ? cconj(o) : o;
                    }
                }
            }
            break;
        }
        propagate_core<C>(psi, tw, last ? p.Hf : p.Hp, q, l, active);
        if (last) {
            // near field: final Ay^-1, natural order store
            if (active)
                for (int g = q; g < R2; g += C::Q) {
                    cpx v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = psi[(R2 * j + g) * S + l];
                    dft8<true>(v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) det[(R2 * j + g) * P + l] = v[j];
                }
        }
    }
}

// ================================================================================================
// Adjoint
// ================================================================================================
template <class C> __global__ void __launch_bounds__(C::NTHREADS) k_multislice_bwd(MsParams p) {
    ADM_SMEM_DECL
    constexpr int P = C::P, R2 = C::R2, S = C::S;
    cpx* psi = (cpx*)smem_raw;          // holds the running adjoint wavefield g
    cpx* ckb = psi + P * S;             // checkpoint staging (cp.async)
    cpx* tw = ckb + P * P;
    const int t = threadIdx.x;
    const bool active = t < C::NACT;
    const int q = t / P, l = t % P;
    const int wf = blockIdx.x;
    const int b = wf / p.M, m = wf % p.M;
    const int py = p.pos[2 * b], px = p.pos[2 * b + 1];
    const int gx = px + l;
    const bool x_in = (gx >= 0) && (gx < p.X);
    const cpx* ckpt = p.ckpt + (size_t)wf * p.Nz * P * P;
    const cpx* gdet = p.det + (size_t)wf * P * P;
    const size_t slice_stride = (size_t)p.Y * p.X;

    init_twiddles<C>(tw, t);
    // prefetch the checkpoint of the last slice
    if (active) {
        for (int g = q; g < R2; g += C::Q)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int y = R2 * j + g;
                async_copy8(&ckb[y * P + l], &ckpt[(size_t)(p.Nz - 1) * P * P + y * P + l], true);
            }
        async_commit();
    }
    __syncthreads();

    // ---- adjoint of the detector transform ---------------------------------------------------------------
    if (p.free_prop == ADM_PROP_FAR) {
        if (active) {
            const int fy = (perm_freq<C>(l) + P / 2) % P;
            for (int g = q; g < C::R1; g += C::Q) {
                cpx v[R2];
#pragma unroll
                for (int k = 0; k < R2; ++k) {
                    int fx = (g + C::R1 * k + P / 2) % P;
                    cpx o = cscale(gdet[fy * P + fx], p.ff_scale);
                    v[k] = p.ff_conj ? cconj(o) : o;
                }
                DftR<R2, true>::run(v);
#pragma unroll
                for (int n = 1; n < R2; ++n) v[n] = cmulc(v[n], tw[g * n]);
#pragma unroll
                for (int n = 0; n < R2; ++n) psi[l * S + R2 * g + n] = v[n];
            }
        }
        __syncthreads();
        if (active) pass_A_inv<C, true>(psi, q, l);
        __syncthreads();
        if (active) pass_B_inv<C, false>(psi, tw, q, l);
        __syncthreads();
    } else if (p.free_prop == ADM_PROP_NEAR) {
        if (active)
            for (int g = q; g < R2; g += C::Q) {
                cpx v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gdet[(R2 * j + g) * P + l];
                dft8<false>(v);
#pragma unroll
                for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw[g * k]);
#pragma unroll
                for (int k = 0; k < 8; ++k) psi[(R2 * k + g) * S + l] = v[k];
            }
        __syncthreads();
        propagate_core<C>(psi, tw, p.Hf, q, l, active);
    }

    for (int z = p.Nz - 1; z >= 0; --z) {
        const bool last = (z == p.Nz - 1);
        // ---- P1': [Ay^-1] . gradient of the modulation . conj(t) . [Ay] ---------------------------------------
        if (active) {
            async_wait_all();
            for (int g = q; g < R2; g += C::Q) {
                cpx v[8], o[8];
                bool inb[8];
                // object voxels of this slice (L2-resident; issued first so the butterfly hides the latency)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    int gy = py + R2 * j + g;
                    inb[j] = x_in && gy >= 0 && gy < p.Y;
                    o[j] = inb[j] ? p.rot_obj[(size_t)z * slice_stride + (size_t)gy * p.X + gx] : mk(0.f, 0.f);
                }
                if (last && p.free_prop == ADM_PROP_NONE) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = gdet[(R2 * j + g) * P + l];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = psi[(R2 * j + g) * S + l];
                    dft8<true>(v);
                    if (last && p.ff_conj) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = cconj(v[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    int y = R2 * j + g, gy = py + y;
                    cpx w = ckb[y * P + l];                       // psi_z (wave entering the slice)
                    if (z > 0) async_copy8(&ckb[y * P + l], &ckpt[(size_t)(z - 1) * P * P + y * P + l], true);
                    cpx tr = transmission(o[j], inb[j], p.unknown_type, p.k1, p.sgn);
                    cpx gt = cmulc(v[j], w);                      // g_t = g' conj(psi)
                    float gd, gb;
                    if (p.unknown_type == ADM_DELTA_BETA) {
                        cpx u = cmulc(gt, tr);                    // g' conj(psi t)
                        gd = -p.sgn * p.k1 * u.y;
                        gb = -p.k1 * u.x;
                    } else {
                        gd = gt.x;
                        gb = gt.y;
                    }
                    if (inb[j]) red_add2(&p.g_obj[(size_t)z * slice_stride + (size_t)gy * p.X + gx], gd, gb);
                    v[j] = cmulc(v[j], tr);                       // g_psi = g' conj(t)
                }
                if (z > 0) {
                    dft8<false>(v);
#pragma unroll
                    for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw[g * k]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) psi[(R2 * k + g) * S + l] = v[k];
                } else if (p.g_probe != nullptr) {
                    if (p.probe_per_pattern) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) p.g_probe[(size_t)wf * P * P + (R2 * j + g) * P + l] = v[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            red_add2(&p.g_probe[(size_t)m * P * P + (R2 * j + g) * P + l], v[j].x, v[j].y);
                    }
                }
            }
            async_commit();
        }
        __syncthreads();
        if (z == 0) break;
        propagate_core<C>(psi, tw, p.Hp, q, l, active);
    }
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
static int fill_params(const adm_multislice_cfg* c, MsParams& p) {
    if (!c) return ADM_ERR_ARG;
    if (c->P != 64 && c->P != 72) return ADM_ERR_UNSUPPORTED;
    if (c->n_slices < 1 || c->B < 1 || c->M < 1 || c->Y < 1 || c->X < 1) return ADM_ERR_ARG;
    if (c->sign_convention != 1 && c->sign_convention != -1) return ADM_ERR_ARG;
    if (c->unknown_type != ADM_DELTA_BETA && c->unknown_type != ADM_REAL_IMAG) return ADM_ERR_ARG;
    if (c->free_prop < ADM_PROP_NONE || c->free_prop > ADM_PROP_NEAR) return ADM_ERR_ARG;
    p.Nz = c->n_slices; p.Y = c->Y; p.X = c->X; p.B = c->B; p.M = c->M;
    p.k1 = c->k1; p.sgn = (float)c->sign_convention;
    p.unknown_type = c->unknown_type; p.free_prop = c->free_prop; p.probe_per_pattern = c->probe_per_pattern;
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

extern "C" int adm_multislice_fwd(const adm_multislice_cfg* cfg, const float* rot_obj, const float* probe, const int* pos,
                                  const float* Hprep, const float* Hfree, float* psi_ckpt, float* det, void* stream) {
    MsParams p;
    int rc = fill_params(cfg, p);
    if (rc) return rc;
    if (!rot_obj || !probe || !pos || !psi_ckpt || !det) return ADM_ERR_ARG;
    if (cfg->n_slices > 1 && !Hprep) return ADM_ERR_ARG;
    if (cfg->free_prop == ADM_PROP_NEAR && !Hfree) return ADM_ERR_ARG;
    p.rot_obj = (const cpx*)rot_obj; p.probe = (const cpx*)probe; p.pos = pos;
    p.Hp = (const cpx*)Hprep; p.Hf = (const cpx*)Hfree;
    p.ckpt = (cpx*)psi_ckpt; p.det = (cpx*)det; p.g_obj = nullptr; p.g_probe = nullptr;
    int nwf = cfg->B * cfg->M;
    if (cfg->P == 72) {
        if (set_smem(k_multislice_fwd<Cfg72>, Cfg72::SMEM)) return ADM_ERR_CUDA;
        ADM_LAUNCH(k_multislice_fwd<Cfg72>, nwf, Cfg72::NTHREADS, Cfg72::SMEM, stream, p);
    } else {
        if (set_smem(k_multislice_fwd<Cfg64>, Cfg64::SMEM)) return ADM_ERR_CUDA;
        ADM_LAUNCH(k_multislice_fwd<Cfg64>, nwf, Cfg64::NTHREADS, Cfg64::SMEM, stream, p);
    }
    return ADM_OK;
}

extern "C" int adm_multislice_bwd(const adm_multislice_cfg* cfg, const float* rot_obj, const int* pos, const float* HprepC,
                                  const float* HfreeC, const float* psi_ckpt, const float* g_det, float* g_rot_obj,
                                  float* g_probe, void* stream) {
    MsParams p;
    int rc = fill_params(cfg, p);
    if (rc) return rc;
    if (!rot_obj || !pos || !psi_ckpt || !g_det || !g_rot_obj) return ADM_ERR_ARG;
    if (cfg->n_slices > 1 && !HprepC) return ADM_ERR_ARG;
    if (cfg->free_prop == ADM_PROP_NEAR && !HfreeC) return ADM_ERR_ARG;
    p.rot_obj = (const cpx*)rot_obj; p.probe = nullptr; p.pos = pos;
    p.Hp = (const cpx*)HprepC; p.Hf = (const cpx*)HfreeC;
    p.ckpt = (cpx*)psi_ckpt; p.det = (cpx*)g_det; p.g_obj = (cpx*)g_rot_obj; p.g_probe = (cpx*)g_probe;
    int nwf = cfg->B * cfg->M;
    if (cfg->P == 72) {
        if (set_smem(k_multislice_bwd<Cfg72>, Cfg72::SMEM)) return ADM_ERR_CUDA;
        ADM_LAUNCH(k_multislice_bwd<Cfg72>, nwf, Cfg72::NTHREADS, Cfg72::SMEM, stream, p);
    } else {
        if (set_smem(k_multislice_bwd<Cfg64>, Cfg64::SMEM)) return ADM_ERR_CUDA;
        ADM_LAUNCH(k_multislice_bwd<Cfg64>, nwf, Cfg64::NTHREADS, Cfg64::SMEM, stream, p);
    }
    return ADM_OK;
}
