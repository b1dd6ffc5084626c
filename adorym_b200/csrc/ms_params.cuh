// Parameter block and point-wise helpers shared by the multislice kernels (multislice.cu, multislice_pfa2.cu).
#pragma once
#include "adm_common.cuh"
#include "../../include/adorym_b200.h"

struct MsParams {
    const cpx* tvol;    // transmission volume T' (adjoint: plain loads of the tile)
    const cpx* probe;
    const int* pos;
    const cpx* Hp;      // prepared in-object transfer function (conj version for the adjoint)
    const cpx* Hf;      // prepared free-space transfer function (near field) or null
    cpx* ckpt;
    cpx* det;           // fwd: out ; bwd: g_det in
    cpx* g_probe;
    int Nz, Y, X, Xp, B, M;
    int free_prop, probe_per_pattern, h_separable;
    float ff_scale;
    int ff_conj;
};

ADM_DEV void red_add2(cpx* addr, float a, float b) {
#ifndef ADM_HOST_EMU
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a), "f"(b) : "memory");
#else
    atomicAdd(&addr->x, a);
    atomicAdd(&addr->y, b);
#endif
}

// psi * (1 + T')
ADM_DEV cpx modulate(cpx v, cpx tp) {
    return mk(fmaf(-v.y, tp.y, fmaf(v.x, tp.x, v.x)), fmaf(v.y, tp.x, fmaf(v.x, tp.y, v.y)));
}
// g * conj(1 + T')
ADM_DEV cpx modulate_conj(cpx v, cpx tp) {
    return mk(fmaf(v.y, tp.y, fmaf(v.x, tp.x, v.x)), fmaf(v.y, tp.x, fmaf(-v.x, tp.y, v.y)));
}


// two-pass prime-factor kernels for P = 72 (multislice_pfa2.cu); selected by adm_ms2_supported()
int adm_ms2_supported(const adm_multislice_cfg* c);
int adm_ms2_fwd_launch(const adm_multislice_cfg* c, const MsParams& p, const float* tvol, void* stream);
int adm_ms2_bwd_launch(const adm_multislice_cfg* c, const MsParams& p, const float* tvol, float* g_tvol, void* stream);

// two wavefields per CTA on packed fp32x2 arithmetic, P = 72 (multislice_x2.cu)
int adm_msx2_fwd_supported(const adm_multislice_cfg* c);
int adm_msx2_fwd_launch(const adm_multislice_cfg* c, const MsParams& p, void* stream);

// register-resident line FFTs, two passes per slice, P = 72 (multislice_col.cu)
int adm_msc_supported(const adm_multislice_cfg* c);
int adm_msc_fwd_launch(const adm_multislice_cfg* c, const MsParams& p, const float* tvol, void* stream);
int adm_msc_bwd_launch(const adm_multislice_cfg* c, const MsParams& p, float* g_tvol, void* stream);
