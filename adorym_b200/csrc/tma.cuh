// TMA (cp.async.bulk[.tensor]) + mbarrier helpers for the multislice kernels (sm_100a), with a host-emulation
// twin (-DADM_HOST_EMU, tests/emu/) in which the copies are performed synchronously by the issuing thread.
//
// Used for three transfers per slice:
//   * forward : the [P][P] tile of the transmission volume T' = T - 1 is fetched by ONE cp.async.bulk.tensor.3d
//               (box P x 2(P+2) floats starting at an even column, out-of-object voxels zero-filled by the TMA unit = vacuum), completion on an mbarrier;
//   * adjoint : the [P][P+2] psi checkpoint of the slice is fetched by ONE cp.async.bulk (contiguous 8 P (P+2) bytes);
//   * adjoint : the [P][P+2] tile of dL/dT is scatter-ADDED into the gradient volume by ONE
//               cp.reduce.async.bulk.tensor.3d (.add, fp32): the L2 performs the reduction sector-wise and the TMA
//               unit clips the overhang, replacing 2 P^2 per-lane RED operations.
#pragma once
#include "adm_common.cuh"

#ifndef ADM_HOST_EMU
#include <cuda.h>
typedef CUtensorMap AdmTileMap;
typedef unsigned long long adm_mbar_t;

ADM_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

ADM_DEV void mbar_init(adm_mbar_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
ADM_DEV void mbar_expect_tx(adm_mbar_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
ADM_DEV void mbar_wait(adm_mbar_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ADM_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ADM_DONE_%=;\n"
        "bra ADM_WAIT_%=;\n"
        "ADM_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// tile load: box [1][P][2P floats] at (x2 = 2*px, y = py, z) -> dense smem tile, zero fill outside the tensor
ADM_DEV void tma_load_tile(void* smem_dst, const AdmTileMap* map, int x2, int y, int z, adm_mbar_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(smem_u32(smem_dst)), "l"((unsigned long long)map), "r"(smem_u32(bar)), "r"(x2), "r"(y), "r"(z)
                 : "memory");
}
// scatter-add a dense smem tile into the tensor (fp32 add performed by the L2), overhang clipped
ADM_DEV void tma_reduce_add_tile(const AdmTileMap* map, int x2, int y, int z, const void* smem_src) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];\n"
                 ::"l"((unsigned long long)map), "r"(smem_u32(smem_src)), "r"(x2), "r"(y), "r"(z) : "memory");
}
ADM_DEV void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, adm_mbar_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
ADM_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// all committed bulk groups have finished READING their shared-memory source
ADM_DEV void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
ADM_DEV void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (before a TMA store / reduce reads them)
ADM_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- host side: tensor-map construction through the driver entry point (no -lcuda link dependency) ----
#include <cuda_runtime.h>
typedef CUresult (*adm_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static adm_encode_fn adm_get_encode() {
    static adm_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (adm_encode_fn)p;
    }
    return fn;
}
// volume [Nz][Y][Xp][2] fp32 (row pitch Xp complex, Xp even), logical extent X <= Xp; box = [P rows][BW complex].
// The TMA unit addresses global memory in 16-byte units: the innermost start coordinate must be an EVEN complex
// index, so the kernels fetch BW = P + 2 columns starting at px & ~1 and index the tile with the offset px & 1.
static int adm_make_tile_map(AdmTileMap* map, const void* base, int Nz, int Y, int X, int Xp, int P, int BW) {
    adm_encode_fn enc = adm_get_encode();
    if (!enc) return 1;
    if ((Xp & 1) || (((uintptr_t)base) & 15)) return 1;
    cuuint64_t dims[3] = {(cuuint64_t)2 * X, (cuuint64_t)Y, (cuuint64_t)Nz};
    cuuint64_t strides[2] = {(cuuint64_t)Xp * 8, (cuuint64_t)Xp * 8 * (cuuint64_t)Y};
    cuuint32_t box[3] = {(cuuint32_t)2 * BW, (cuuint32_t)P, 1};
    if ((BW & 1) || 2 * BW > 256) return 1;
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}
#define ADM_GRID_CONSTANT __grid_constant__

#else   // ------------------------------- host emulation -------------------------------------------------
#include <atomic>
struct AdmTileMap {
    float* base;
    int Nz, Y, X, Xp, P, BW;
};
typedef unsigned long long adm_mbar_t;    // low 32 bits: number of completed phases

static inline std::atomic_ref<unsigned long long> mbar_ref(adm_mbar_t* b) { return std::atomic_ref<unsigned long long>(*b); }
static inline void mbar_init(adm_mbar_t* bar, unsigned) { mbar_ref(bar).store(0); }
static inline void mbar_expect_tx(adm_mbar_t*, unsigned) {}
static inline void mbar_wait(adm_mbar_t* bar, unsigned parity) {
    for (;;) {
        unsigned long long v = mbar_ref(bar).load();
        if ((v & 1ull) != (unsigned long long)parity) break;
        mbar_ref(bar).wait(v);
    }
}
static inline void tma_load_tile(void* dst, const AdmTileMap* m, int x2, int y, int z, adm_mbar_t* bar) {
    float* d = (float*)dst;
    if (x2 & 3) abort();    // the hardware raises "illegal instruction" for a start that is not 16-byte aligned
    for (int r = 0; r < m->P; ++r)
        for (int c = 0; c < 2 * m->BW; ++c) {
            int yy = y + r, xx = x2 + c;
            bool in = z >= 0 && z < m->Nz && yy >= 0 && yy < m->Y && xx >= 0 && xx < 2 * m->X;
            d[r * 2 * m->BW + c] = in ? m->base[((size_t)z * m->Y + yy) * 2 * m->Xp + xx] : 0.f;
        }
    mbar_ref(bar).fetch_add(1);
    mbar_ref(bar).notify_all();
}
static inline void tma_reduce_add_tile(const AdmTileMap* m, int x2, int y, int z, const void* src) {
    const float* s = (const float*)src;
    if ((x2 & 3) || x2 < 0 || y < 0 || z < 0) abort();   // hardware: negative start coordinates of a TMA store / reduce fault
    for (int r = 0; r < m->P; ++r)
        for (int c = 0; c < 2 * m->BW; ++c) {
            int yy = y + r, xx = x2 + c;
            bool in = z >= 0 && z < m->Nz && yy >= 0 && yy < m->Y && xx >= 0 && xx < 2 * m->X;
            if (in) atomicAdd(&m->base[((size_t)z * m->Y + yy) * 2 * m->Xp + xx], s[r * 2 * m->BW + c]);
        }
}
static inline void bulk_load(void* dst, const void* src, unsigned bytes, adm_mbar_t* bar) {
    memcpy(dst, src, bytes);
    mbar_ref(bar).fetch_add(1);
    mbar_ref(bar).notify_all();
}
static inline void bulk_commit() {}
static inline void bulk_wait_read_all() {}
static inline void bulk_wait_all() {}
static inline void fence_proxy_async() {}
static int adm_make_tile_map(AdmTileMap* map, const void* base, int Nz, int Y, int X, int Xp, int P, int BW) {
    map->base = (float*)base; map->Nz = Nz; map->Y = Y; map->X = X; map->Xp = Xp; map->P = P; map->BW = BW;
    return ((Xp & 1) || (BW & 1)) ? 1 : 0;
}
#define ADM_GRID_CONSTANT
#endif
