// TEST INFRASTRUCTURE ONLY: a minimal "CUDA on the CPU" execution layer so that the kernel sources in
// adorym_b200/csrc can be compiled as C++ (-DADM_HOST_EMU) and their LOGIC unit-tested in the GPU-less
// build container.  One OS thread per CUDA thread, CTAs run one after another.  The product never
// loads the library built from this header (adorym_b200/_lib.py only opens the nvcc-built .so).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct emu_dim3 { unsigned x, y, z; emu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef emu_dim3 dim3;
typedef void* cudaStream_t;
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

struct EmuBarrier {
    std::mutex m; std::condition_variable cv; unsigned n, count = 0, gen = 0;
    explicit EmuBarrier(unsigned n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        unsigned g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

// named barriers (bar.sync id, count): created on first use, one set per CTA
struct EmuNamedBarriers {
    std::mutex m;
    std::vector<std::pair<int, EmuBarrier*>> bars;
    EmuBarrier* get(int id, unsigned n) {
        std::lock_guard<std::mutex> lk(m);
        for (auto& b : bars) if (b.first == id) return b.second;
        bars.emplace_back(id, new EmuBarrier(n));
        return bars.back().second;
    }
    ~EmuNamedBarriers() { for (auto& b : bars) delete b.second; }
};
static thread_local EmuNamedBarriers* emu_named = nullptr;
static inline void emu_named_sync(int id, unsigned n);

static thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
static thread_local EmuBarrier* emu_bar = nullptr;
static thread_local char* emu_smem = nullptr;

#define __syncthreads() emu_bar->wait()
static inline void emu_named_sync(int id, unsigned n) { emu_named->get(id, n)->wait(); }
#define __threadfence() std::atomic_thread_fence(std::memory_order_seq_cst)
#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

static inline float emu_atomic_add(float* p, float v) {
    std::atomic_ref<float> r(*p);
    float old = r.load();
    while (!r.compare_exchange_weak(old, old + v)) {}
    return old;
}
#define atomicAdd(p, v) emu_atomic_add((p), (v))

template <class F>
static void emu_launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
    unsigned nthreads = block.x * block.y * block.z;
    std::vector<char> smem(smem_bytes + 64);
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        EmuBarrier bar(nthreads);
        EmuNamedBarriers named;
        std::memset(smem.data(), 0xCD, smem.size());   // poison: uninitialised smem reads show up as garbage
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t) {
            th.emplace_back([&, t] {
                threadIdx = emu_dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                blockIdx = emu_dim3(bx, by, bz);
                blockDim = block; gridDim = grid;
                emu_bar = &bar;
                emu_named = &named;
                emu_smem = smem.data();
                body();
            });
        }
        for (auto& x : th) x.join();
    }
}
