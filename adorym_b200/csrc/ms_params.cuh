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
    cpx* g_ckpt;        // adjoint only, optional: dL/d psi_z (gradient w.r.t. the wave entering slice z) [B][M][Nz][P][P]
    int Nz, Y, X, Xp, B, M;
    int free_prop, probe_per_pattern, h_separable, h_per_slice, hfree_separable;
    float ff_scale;
    int ff_conj;
    // persistent balanced schedule (multislice.cu): the nwf * Nz slice steps of a launch are laid end to end and cut in
    // gridDim.x intervals of T steps; a wavefield cut by an interval boundary is handed from one CTA to the next through
    // a raw copy of the shared-memory wavefield (handoff slot) and a flag.  split = 0: one wavefield per CTA.
    int split, T;
    int* flags;          // [slots] hand-off flags, 0 outside a launch
    float4* handoff;     // [slots][slot_f4] raw copies of the psi buffer
    int slot_f4;
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


