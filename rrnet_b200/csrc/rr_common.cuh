// Shared helpers for librrnet_b200 (sm_100a).  Internal header, not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#include "../../include/rrnet_b200.h"

#define RR_API extern "C" __attribute__((visibility("default")))

namespace rr {

extern std::atomic<uint64_t> g_launches;   // defined in rr_api.cu

// Count one kernel launch and fold its launch status into `rc` (first error wins).
#define RR_LAUNCHED(rc)                                                    \
    do {                                                                   \
        ::rr::g_launches.fetch_add(1, std::memory_order_relaxed);          \
        cudaError_t _e = cudaPeekAtLastError();                            \
        if (_e != cudaSuccess && (rc) == 0) (rc) = (int)_e;                \
    } while (0)

// Optional per-kernel timing (rr_kernel_trace_begin / _end): the calling thread hands in CUDA events and every
// launch it issues afterwards records the next one on the launch's stream, so that event[i] - event[i-1] is the
// device time of kernel i when the queue never runs dry.  Off (events == nullptr) unless a bench asks for it.
struct KernelTrace { void* const* events; const char** names; int cap; int n; };
extern thread_local KernelTrace g_ktrace;   // defined in rr_api.cu
inline void ktrace_mark(const char* name, cudaStream_t st) {
    KernelTrace& t = g_ktrace;
    if (t.events && t.n < t.cap) {
        if (t.names) t.names[t.n] = name;
        (void)cudaEventRecord((cudaEvent_t)t.events[t.n++], st);
    }
}
#define RR_LAUNCHED_K(rc, name, st)                                        \
    do {                                                                   \
        RR_LAUNCHED(rc);                                                   \
        ::rr::ktrace_mark(name, st);                                       \
    } while (0)

#define RR_CUDA(call, rc)                                                  \
    do {                                                                   \
        cudaError_t _e = (call);                                           \
        if (_e != cudaSuccess && (rc) == 0) (rc) = (int)_e;                \
    } while (0)

// Programmatic dependent launch (PDL) along the eval chain: a kernel launched with launch_pdl() may be scheduled while
// its predecessor on the stream is still running; its threads block in pdl_wait() - the first statement of every such
// kernel - until the predecessor grid has completed and its writes are visible.  pdl_wait() is a no-op in a kernel that
// was launched the ordinary way, and pdl_trigger() (called right after it) lets the NEXT kernel's blocks be staged as
// soon as all blocks of this grid are resident, so the launch latency between the ~14 short kernels of a step overlaps
// with their execution.  Captured into CUDA graphs as programmatic dependency edges.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define RR_PDL_PROLOGUE() do { ::rr::pdl_wait(); ::rr::pdl_trigger(); } while (0)

extern int g_pdl_enabled;                  // rr_set_pdl (default 1), defined in rr_api.cu
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl_enabled ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // the launch status is picked up by RR_LAUNCHED
}

constexpr int kSMs = 148;   // B200
// SMs the persistent kernels leave free (rr_set_sm_reserve).  Per calling host thread: the value is read when a
// launch is issued (or captured into a graph) by that thread, so two threads driving two streams do not interfere.
extern thread_local int g_sm_reserve;      // defined in rr_api.cu
inline int sms_for_persistent() { const int r = g_sm_reserve; return kSMs - (r < 0 ? 0 : (r > kSMs - 1 ? kSMs - 1 : r)); }

// Function attributes (e.g. the dynamic shared memory opt-in) are per device: do `first-use` work once per device
// of the process.  Two racing threads may both do it, which is harmless (the calls are idempotent).
struct OncePerDevice {
    std::atomic<uint64_t> done[4]{};          // up to 256 devices
    bool need(int* dev_out) {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 256) { *dev_out = -1; return true; }
        *dev_out = d;
        return !(done[d >> 6].load(std::memory_order_acquire) & (1ull << (d & 63)));
    }
    void mark(int d) { if (d >= 0) done[d >> 6].fetch_or(1ull << (d & 63), std::memory_order_release); }
};

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Carves a caller-provided workspace into 256-byte aligned pieces.
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(reinterpret_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + off);
        off = align_up(off + count * sizeof(T));
        return p;
    }
};

// Monotone fp32 -> uint32 ordering key (larger float <=> larger key).
__host__ __device__ __forceinline__ uint32_t f2key(float v) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(v);
#else
    uint32_t u;
    memcpy(&u, &v, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ float sigmoid_f32(float z) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-z)));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load that does not pollute L1 (data is touched once per CTA)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

}  // namespace rr
