// Hardware probe (run on the GPU box): which coordinates / box shapes does cp.reduce.async.bulk.tensor accept?
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cstdio>
#include <vector>
#include "../adorym_b200/csrc/tma.cuh"

__global__ void k_probe(const __grid_constant__ CUtensorMap map, int x2, int y, int z, int nfl, int do_load) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* tile = (float*)smem;
    adm_mbar_t* bar = (adm_mbar_t*)(smem + 65536);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (do_load) {
        if (threadIdx.x == 0) { mbar_expect_tx(bar, nfl * 4); tma_load_tile(tile, &map, x2, y, z, bar); }
        mbar_wait(bar, 0);
        __syncthreads();
        return;
    }
    for (int i = threadIdx.x; i < nfl; i += blockDim.x) tile[i] = 1.0f;
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) { tma_reduce_add_tile(&map, x2, y, z, tile); bulk_commit(); bulk_wait_all(); }
}

int main() {
    const int Nz = 4, Y = 100, X = 100, Xp = 100;
    float* d; cudaMalloc(&d, sizeof(float) * 2 * Nz * Y * Xp);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    int Ps[2] = {64, 72};
    int coords[][3] = {{0, 0, 0}, {8, 4, 1}, {100, 50, 1}, {-8, 0, 0}, {0, -5, 0}, {-72, -36, 0}, {180, 90, 3}};
    for (int do_load = 0; do_load < 2; ++do_load)
    for (int ip = 0; ip < 2; ++ip) {
        int P = Ps[ip], BW = P + 2;
        CUtensorMap map;
        if (adm_make_tile_map(&map, d, Nz, Y, X, Xp, P, BW)) { printf("map failed\n"); return 1; }
        for (auto& c : coords) {
            cudaMemset(d, 0, sizeof(float) * 2 * Nz * Y * Xp);
            k_probe<<<1, 128, 65536 + 64>>>(map, c[0], c[1], c[2], 2 * BW * P, do_load);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> h(2 * Nz * Y * Xp);
            double sum = 0;
            if (e == cudaSuccess) { cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost); for (float v : h) sum += v; }
            printf("%s P=%d x2=%d y=%d z=%d -> %s sum=%.0f\n", do_load ? "load" : "reduce", P, c[0], c[1], c[2], cudaGetErrorString(e), sum);
            if (e != cudaSuccess) { printf("context dead, stopping\n"); return 0; }
        }
    }
    return 0;
}
