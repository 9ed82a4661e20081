// emu_cuda.hpp -- a small SIMT emulator: runs the product's CUDA kernels (linevis_b200/csrc/*.cuh, unchanged source) on the HOST.
//
// TEST INFRASTRUCTURE ONLY (CPU test suite; nothing here is linked into the product).  tests/emu/build_emu.py compiles
// linevis_b200/csrc/lv_api.cu with plain g++ against this header: every CUDA thread becomes a fiber (ucontext), a block is
// 32..1024 fibers scheduled in lockstep by one OS thread, blocks are spread over OS threads with OpenMP.  Warp collectives
// (__ballot_sync, __shfl_*_sync, __syncwarp, __match_any_sync, __activemask) and __syncthreads are rendez-vous points of the
// scheduler; __shared__ becomes per-OS-thread static storage; atomics map to GCC __atomic builtins; the CUDA runtime calls the
// host code makes (cudaMalloc, cudaMemcpyAsync, events, ...) are implemented on host memory (emu_cudart.inc).  With strict
// float flags (-ffp-contract=off, FMA only where the source says fmaf) the emulated library reproduces the GPU results bit for
// bit, so the whole C ABI -- BVH build, packet traversal, the AO ray stream, PPLL gather / resolve -- can be checked against the
// oracle without a GPU.  It is slow (a collective costs 64 fiber switches): tiny frames only.
#pragma once
#define LV_HOST_EMU_SIMT 1   // product headers: compile the warp-collective code too
#include <cuda_runtime.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <vector>

// ---- qualifiers: cuda_runtime.h (host_defines.h) turns them into ignored attributes; the ones that matter are redefined
#undef __shared__
#define __shared__ static thread_local
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
#undef __constant__
#define __constant__
#undef __device__
#define __device__
#undef __global__
#define __global__
#undef __host__
#define __host__

namespace emu {

enum State : int { RUN = 0, W_BALLOT, W_SHFL, W_SHFL_XOR, W_SHFL_UP, W_SYNCWARP, W_MATCH, W_ACTIVE, W_BLOCK, DONE };

struct Lane {
    ucontext_t ctx;
    int state = DONE;
    unsigned mask = 0;
    uint64_t val = 0, result = 0;
    int arg = 0;
    uint3 tid;
};

struct Block {
    std::vector<Lane> lanes;
    char* stacks = nullptr;   // fiber stacks: one uninitialised, reused allocation per OS thread (stack_pool)
    ucontext_t sched;
    int cur = 0, nthreads = 0;
    uint3 bidx, bdim, gdim;
    std::function<void()> body;
    std::vector<char> dyn;
};

inline Block*& cur_block() { static thread_local Block* b = nullptr; return b; }
inline Lane& cur_lane() { Block* b = cur_block(); return b->lanes[b->cur]; }
inline void* dyn_smem() { return cur_block()->dyn.data(); }

constexpr size_t kStackBytes = 128 * 1024;
inline char* stack_pool(size_t bytes) {
    static thread_local char* pool = nullptr;
    static thread_local size_t cap = 0;
    if (bytes > cap) { std::free(pool); pool = static_cast<char*>(std::malloc(bytes)); cap = bytes; }
    return pool;
}

inline void lane_entry() {
    Block* b = cur_block();
    b->body();
    b->lanes[b->cur].state = DONE;   // falls back into the scheduler through uc_link
}

inline uint64_t wait(State kind, unsigned mask, uint64_t val, int arg) {
    Block* b = cur_block();
    Lane& l = b->lanes[b->cur];
    l.state = kind; l.mask = mask; l.val = val; l.arg = arg;
    swapcontext(&l.ctx, &b->sched);
    return l.result;
}

// resolve the warp-level rendez-vous of warp w; true if some lane was released
inline bool resolve_warp(Block* b, int w) {
    const int base = w * 32, n = std::min(32, b->nthreads - base);
    bool released = false;
    for (int kind = W_BALLOT; kind <= W_ACTIVE; kind++) {
        unsigned here = 0;
        for (int i = 0; i < n; i++) if (b->lanes[base + i].state == kind) here |= 1u << i;
        if (!here) continue;
        // participants: the lanes named by the (common) mask that have not exited; W_ACTIVE: whoever is here
        unsigned want = 0;
        if (kind == W_ACTIVE) want = here;
        else {
            unsigned m = 0;
            for (int i = 0; i < n; i++) if (here >> i & 1u) m |= b->lanes[base + i].mask;
            for (int i = 0; i < n; i++) if ((m >> i & 1u) && b->lanes[base + i].state != DONE) want |= 1u << i;
        }
        if ((want & ~here) != 0u) continue;   // somebody named by the mask is still elsewhere: not yet
        for (int i = 0; i < n; i++) {
            if (!(here >> i & 1u)) continue;
            Lane& l = b->lanes[base + i];
            switch (kind) {
                case W_BALLOT: { unsigned r = 0; for (int j = 0; j < n; j++) if ((here >> j & 1u) && (l.mask >> j & 1u) && b->lanes[base + j].val) r |= 1u << j; l.result = r; break; }
                case W_SHFL: { int s = l.arg & 31; l.result = (s < n && (here >> s & 1u)) ? b->lanes[base + s].val : l.val; break; }
                case W_SHFL_XOR: { int s = i ^ l.arg; l.result = (s < n && (here >> s & 1u)) ? b->lanes[base + s].val : l.val; break; }
                case W_SHFL_UP: { int s = i - l.arg; l.result = (s >= 0 && (here >> s & 1u)) ? b->lanes[base + s].val : l.val; break; }
                case W_MATCH: { unsigned r = 0; for (int j = 0; j < n; j++) if ((here >> j & 1u) && (l.mask >> j & 1u) && b->lanes[base + j].val == l.val) r |= 1u << j; l.result = r; break; }
                case W_ACTIVE: l.result = here; break;
                default: l.result = 0; break;
            }
        }
        for (int i = 0; i < n; i++) if (here >> i & 1u) b->lanes[base + i].state = RUN;
        released = true;
    }
    return released;
}

// LV_EMU_ORDER=reverse|random: the order in which the scheduler runs the lanes of a warp (and the warps of a block) between two
// rendez-vous points.  Results must not depend on it: a kernel that does has a race the hardware's lockstep execution may hide
// (e.g. a missing __syncwarp between a shared-memory write and another lane's read).  Default: ascending.
inline int order_mode() {
    static const int m = [] { const char* e = std::getenv("LV_EMU_ORDER"); return !e ? 0 : (std::strcmp(e, "reverse") == 0 ? 1 : (std::strcmp(e, "random") == 0 ? 2 : 0)); }();
    return m;
}
inline void make_order(int* perm, int n) {
    static thread_local uint64_t rng = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < n; i++) perm[i] = order_mode() == 1 ? n - 1 - i : i;
    if (order_mode() == 2)
        for (int i = n - 1; i > 0; i--) {
            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
            std::swap(perm[i], perm[int(rng % uint64_t(i + 1))]);
        }
}

inline void run_block(Block* b) {
    cur_block() = b;
    for (int t = 0; t < b->nthreads; t++) {
        Lane& l = b->lanes[t];
        getcontext(&l.ctx);
        l.ctx.uc_stack.ss_sp = b->stacks + size_t(t) * kStackBytes;
        l.ctx.uc_stack.ss_size = kStackBytes;
        l.ctx.uc_link = &b->sched;
        makecontext(&l.ctx, (void (*)())lane_entry, 0);
        l.state = RUN;
        l.tid = make_uint3(unsigned(t) % b->bdim.x, (unsigned(t) / b->bdim.x) % b->bdim.y, unsigned(t) / (b->bdim.x * b->bdim.y));
    }
    const int nwarps = (b->nthreads + 31) / 32;
    while (true) {
        bool progress = false, all_done = true;
        int worder[64];
        if (nwarps <= 64) make_order(worder, nwarps);
        for (int wk = 0; wk < nwarps; wk++) {
            const int w = nwarps <= 64 ? worder[wk] : wk;
            bool again = true;
            while (again) {
                again = false;
                const int base = w * 32, n = std::min(32, b->nthreads - base);
                int lorder[32];
                make_order(lorder, n);
                for (int k = 0; k < n; k++) {
                    const int i = lorder[k];
                    Lane& l = b->lanes[base + i];
                    while (l.state == RUN) { b->cur = base + i; swapcontext(&b->sched, &l.ctx); progress = true; }
                }
                if (resolve_warp(b, w)) { again = true; progress = true; }
            }
        }
        bool all_block = true;
        for (int t = 0; t < b->nthreads; t++) {
            const int s = b->lanes[t].state;
            if (s != DONE) all_done = false;
            if (s != DONE && s != W_BLOCK) all_block = false;
        }
        if (all_done) break;
        if (all_block) { for (int t = 0; t < b->nthreads; t++) if (b->lanes[t].state == W_BLOCK) b->lanes[t].state = RUN; progress = true; }
        if (!progress) {
            std::fprintf(stderr, "emu: deadlock in block (%u,%u,%u): lane states", b->bidx.x, b->bidx.y, b->bidx.z);
            for (int t = 0; t < std::min(b->nthreads, 64); t++) std::fprintf(stderr, " %d", b->lanes[t].state);
            std::fprintf(stderr, "\n");
            std::abort();
        }
    }
    cur_block() = nullptr;
}

template <class... P>
struct Launcher {
    void (*kernel)(P...);
    dim3 grid, block;
    size_t smem;
    template <class... A>
    void operator()(A&&... a) const {
        std::tuple<std::decay_t<P>...> args(std::forward<A>(a)...);
        const long long nblocks = (long long)grid.x * grid.y * grid.z;
        const int nthreads = int(block.x * block.y * block.z);
        auto k = kernel;
#pragma omp parallel
        {
            Block b;
            b.nthreads = nthreads; b.bdim = block; b.gdim = grid;
            b.lanes.resize(nthreads);
            b.stacks = stack_pool(size_t(nthreads) * kStackBytes);
            b.dyn.resize(smem + 16);
            b.body = [&]() { std::apply(k, args); };
#pragma omp for schedule(dynamic, 1)
            for (long long i = 0; i < nblocks; i++) {
                b.bidx = make_uint3(unsigned(i % grid.x), unsigned((i / grid.x) % grid.y), unsigned(i / ((long long)grid.x * grid.y)));
                run_block(&b);
            }
        }
    }
};
template <class... P>
Launcher<P...> launch(void (*k)(P...), dim3 g, dim3 b, size_t smem, cudaStream_t) { return Launcher<P...>{k, g, b, smem}; }

template <class T> inline uint64_t bits_of(T v) { uint64_t r = 0; std::memcpy(&r, &v, sizeof(T)); return r; }
template <class T> inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace emu

#define threadIdx (emu::cur_lane().tid)
#define blockIdx (emu::cur_block()->bidx)
#define blockDim (emu::cur_block()->bdim)
#define gridDim (emu::cur_block()->gdim)
#define EMU_LAUNCH(kernel, g, b, s, st) emu::launch(kernel, dim3(g), dim3(b), size_t(s), st)

// ---- warp / block collectives
inline unsigned __ballot_sync(unsigned mask, int pred) { return unsigned(emu::wait(emu::W_BALLOT, mask, pred ? 1 : 0, 0)); }
template <class T> inline T __shfl_sync(unsigned mask, T v, int src) { return emu::from_bits<T>(emu::wait(emu::W_SHFL, mask, emu::bits_of(v), src)); }
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int d) { return emu::from_bits<T>(emu::wait(emu::W_SHFL_XOR, mask, emu::bits_of(v), d)); }
template <class T> inline T __shfl_up_sync(unsigned mask, T v, int d) { return emu::from_bits<T>(emu::wait(emu::W_SHFL_UP, mask, emu::bits_of(v), d)); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::wait(emu::W_SYNCWARP, mask, 0, 0); }
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) { return unsigned(emu::wait(emu::W_MATCH, mask, emu::bits_of(v), 0)); }
inline unsigned __activemask() { return unsigned(emu::wait(emu::W_ACTIVE, 0, 0, 0)); }
inline void __syncthreads() { emu::wait(emu::W_BLOCK, 0, 0, 0); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- bit / conversion intrinsics (the float <-> uint ones come from lv_math.cuh's LV_HOST_EMU block)
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs(int(v)); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(unsigned(v)); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }

// ---- atomics (blocks run on different OS threads)
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicSub(int* p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T emu_atomic_minmax(T* p, T v, bool is_min) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while ((is_min ? v < old : v > old) && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int atomicMin(int* p, int v) { return emu_atomic_minmax(p, v, true); }
inline unsigned atomicMin(unsigned* p, unsigned v) { return emu_atomic_minmax(p, v, true); }
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { return emu_atomic_minmax(p, v, true); }
inline int atomicMax(int* p, int v) { return emu_atomic_minmax(p, v, false); }
inline unsigned atomicMax(unsigned* p, unsigned v) { return emu_atomic_minmax(p, v, false); }

// (integer min / max overloads: lv_math.cuh's LV_HOST_EMU block)

// ---- cub::DeviceRadixSort::SortPairs (BVH builder; a stable sort on the selected key bits) and cub::DeviceScan::ExclusiveSum (PPLL fill)
namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out, int n,
                                 int begin_bit, int end_bit, cudaStream_t) {
        if (!tmp) { tmp_bytes = 16; return cudaSuccess; }
        std::vector<int> order(n);
        for (int i = 0; i < n; i++) order[i] = i;
        const K mask = (end_bit - begin_bit >= int(sizeof(K) * 8)) ? ~K(0) : (((K(1) << (end_bit - begin_bit)) - 1) << begin_bit);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (keys_in[a] & mask) < (keys_in[b] & mask); });
        for (int i = 0; i < n; i++) { keys_out[i] = keys_in[order[i]]; vals_out[i] = vals_in[order[i]]; }
        return cudaSuccess;
    }
};
struct DeviceScan {
    template <class In, class Out>
    static cudaError_t ExclusiveSum(void* tmp, size_t& tmp_bytes, const In* in, Out* out, int n, cudaStream_t) {
        if (!tmp) { tmp_bytes = 16; return cudaSuccess; }
        Out acc = 0;
        for (int i = 0; i < n; i++) { const Out v = Out(in[i]); out[i] = acc; acc += v; }
        return cudaSuccess;
    }
};
}  // namespace cub

// cudaFuncSetAttribute(kernel, ...): cuda_runtime.h only declares the kernel-pointer template for nvcc
template <class... P> inline cudaError_t cudaFuncSetAttribute(void (*)(P...), cudaFuncAttribute, int) { return cudaSuccess; }
