// cuda_runtime.h (SIMT shim) — development aid, NOT product and NOT a fallback.
//
// Lets the kernel sources of multilanguagekdtree_b200/csrc/*.cu be compiled by g++ and EXECUTED ON THE CPU, one fiber
// per CUDA thread, so that kernel logic (control flow, shared-memory hand-offs, warp votes, look-back protocols) can be
// debugged with gdb / asan / printf and checked against the oracle without spending GPU time.  It is found instead of
// the real <cuda_runtime.h> only when this directory is put first on the include path (tests/devtools/simt/build.py);
// the result, libkdnb_simt.so, exports the same C ABI and is loaded only by tests/devtools/simt/run.py.
// Nothing under multilanguagekdtree_b200/ refers to it, nothing here is timed or shipped.
//
// Model: a kernel launch runs its CTAs on a few OS threads (CTAs are handed out in index order, so a CTA only ever
// waits for CTAs that have started — the same guarantee the look-back kernels rely on); inside a CTA every CUDA thread
// is a fiber (own stack, cooperative switch), __syncthreads and the *_sync warp collectives are yield points that
// complete when every live participant has arrived.  __shared__ is `static thread_local` (one CTA per OS thread at a
// time).  Floating point: g++ -ffp-contract=off, fma() is the exact libm/hardware fma; rsqrt.approx is modelled as
// 1/sqrt truncated to the high word (like MUFU.RSQ64H it is only an estimate; results that depend on it are compared
// with tolerances, never bit for bit, against the GPU).
#pragma once
#define KDNB_SIMT 1

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <functional>
#include <tuple>
#include <utility>

// ---------------------------------------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static thread_local
#define __constant__ static const
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

// ---------------------------------------------------------------------------------------------- vector types
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) double4 { double x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---------------------------------------------------------------------------------------------- the SIMT machine
namespace simt {
struct Fiber {
  void* sp;
  uint3 tid;
  int lane, warp;
  bool done;
};
struct Warp {
  uint32_t live, arrived, gen;
  uint64_t vals[2][32];
};
struct Block {
  uint3 bid;
  dim3 bdim, gdim;
  int nthreads, nlive;
  uint32_t bar_arrived, bar_gen;
  Fiber* fibers;
  Warp* warps;
  void* sched_sp;
  unsigned char* dyn_smem;
  const std::function<void()>* body;
};
extern thread_local Block* B;
extern thread_local Fiber* cur;
void yield();
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
unsigned char* dyn_smem();

// warp-wide exchange: every participating lane deposits v, gets the snapshot of all lanes' values
const uint64_t* warp_exchange(uint32_t mask, uint64_t v);
void block_barrier();

static inline double rsqrt_approx(double x) {  // model of MUFU.RSQ64H + zero low word
  double r = 1.0 / sqrt(x);
  uint64_t u;
  memcpy(&u, &r, 8);
  u &= 0xffffffff00000000ull;
  memcpy(&r, &u, 8);
  return r;
}
template <typename T>
static inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t u = 0;
  memcpy(&u, &v, sizeof(T));
  return u;
}
template <typename T>
static inline T from_bits(uint64_t u) {
  T v;
  memcpy(&v, &u, sizeof(T));
  return v;
}
}  // namespace simt

// built-in variables: plain thread-local PODs, set by the scheduler (threadIdx on every switch to a fiber)
extern thread_local uint3 threadIdx, blockIdx, blockDim, gridDim;

static inline void __syncthreads() { simt::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_exchange(mask, 0); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  const uint64_t* v = simt::warp_exchange(mask, pred ? 1 : 0);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l)
    if (((mask >> l) & 1u) && v[l]) r |= 1u << l;
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return (__ballot_sync(mask, pred) & mask) == mask; }
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src) {
  return simt::from_bits<T>(simt::warp_exchange(mask, simt::to_bits(v))[src & 31]);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int x) {
  return simt::from_bits<T>(simt::warp_exchange(mask, simt::to_bits(v))[(simt::cur->lane ^ x) & 31]);
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned d) {
  const uint64_t* s = simt::warp_exchange(mask, simt::to_bits(v));
  const int l = simt::cur->lane;
  return l >= (int)d ? simt::from_bits<T>(s[l - d]) : v;
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned d) {
  const uint64_t* s = simt::warp_exchange(mask, simt::to_bits(v));
  const int l = simt::cur->lane;
  return l + (int)d < 32 ? simt::from_bits<T>(s[l + d]) : v;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
  const uint64_t* s = simt::warp_exchange(mask, v);
  unsigned r = 0xffffffffu;
  for (int l = 0; l < 32; ++l)
    if (((mask >> l) & 1u) && (unsigned)s[l] < r) r = (unsigned)s[l];
  return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  const uint64_t* s = simt::warp_exchange(mask, v);
  unsigned r = 0u;
  for (int l = 0; l < 32; ++l)
    if (((mask >> l) & 1u) && (unsigned)s[l] > r) r = (unsigned)s[l];
  return r;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
  const uint64_t* s = simt::warp_exchange(mask, v);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l)
    if (((mask >> l) & 1u) && (unsigned)s[l] == v) r |= 1u << l;
  return r;
}

// ---------------------------------------------------------------------------------------------- atomics, fences
template <typename T>
static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return o;
}
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return o;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline long long clock64() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }

// ---------------------------------------------------------------------------------------------- math / bit intrinsics
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline long long __double_as_longlong(double x) { return simt::from_bits<long long>(simt::to_bits(x)); }
static inline double __longlong_as_double(long long x) { return simt::from_bits<double>((uint64_t)x); }
static inline int __double2hiint(double x) { return (int)(simt::to_bits(x) >> 32); }
static inline int __double2loint(double x) { return (int)(simt::to_bits(x) & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo) {
  return simt::from_bits<double>(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
}
static inline unsigned __double2uint_rz(double x) {  // saturating, NaN -> 0
  if (!(x > 0.0)) return 0u;
  if (x >= 4294967295.0) return 0xffffffffu;
  return (unsigned)x;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------- runtime API (host side)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
static inline const char* cudaGetErrorString(cudaError_t e) {
  return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "not supported by the SIMT shim");
}
typedef struct simt_stream* cudaStream_t;
struct simt_event { std::chrono::steady_clock::time_point t; };
typedef simt_event* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
typedef void* cudaGraphNode_t;
typedef unsigned long long cudaGraphConditionalHandle;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeThreadLocal = 1 };
enum { cudaStreamSetCaptureDependencies = 1 };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaGraphCondAssignDefault = 1 };
enum cudaGraphNodeType { cudaGraphNodeTypeConditional = 13 };
enum cudaGraphConditionalNodeType { cudaGraphCondTypeIf = 0 };
struct cudaConditionalNodeParams {
  cudaGraphConditionalHandle handle;
  cudaGraphConditionalNodeType type;
  unsigned size;
  cudaGraph_t* phGraph_out;
};
struct cudaGraphNodeParams {
  cudaGraphNodeType type;
  cudaConditionalNodeParams conditional;
};
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute {
  cudaLaunchAttributeID id;
  cudaLaunchAttributeValue val;
};
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes;
  cudaStream_t stream;
  cudaLaunchAttribute* attrs;
  unsigned numAttrs;
};

template <typename T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  void* q = nullptr;
  if (posix_memalign(&q, 256, bytes ? bytes : 1) != 0) return cudaErrorMemoryAllocation;
  memset(q, 0xA5, bytes);  // device memory is not zero-initialised: make reads of unwritten data visible
  *p = (T*)q;
  return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return posix_memalign(p, 256, bytes ? bytes : 1) ? cudaErrorMemoryAllocation : cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)(uintptr_t)16; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* st) { *st = cudaStreamCaptureStatusNone; return cudaSuccess; }
static inline cudaError_t cudaStreamGetCaptureInfo(cudaStream_t, cudaStreamCaptureStatus* st, unsigned long long*, cudaGraph_t*, const cudaGraphNode_t**, size_t*) {
  *st = cudaStreamCaptureStatusNone;
  return cudaSuccess;
}
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
static inline cudaError_t cudaStreamBeginCaptureToGraph(cudaStream_t, cudaGraph_t, const cudaGraphNode_t*, const void*, size_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
static inline cudaError_t cudaStreamUpdateCaptureDependencies(cudaStream_t, cudaGraphNode_t*, size_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphConditionalHandleCreate(cudaGraphConditionalHandle*, cudaGraph_t, unsigned, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphAddNode(cudaGraphNode_t*, cudaGraph_t, const cudaGraphNode_t*, size_t, cudaGraphNodeParams*) { return cudaErrorNotSupported; }
static inline void cudaGraphSetConditional(cudaGraphConditionalHandle, unsigned) {}
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new simt_event(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
enum { cudaEventRecordDefault = 0, cudaEventRecordExternal = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }  // (launches run to completion)
static inline cudaError_t cudaEventRecordWithFlags(cudaEvent_t e, cudaStream_t s, unsigned) { return cudaEventRecord(e, s); }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
// "peer" contexts of the multi-rank model are threads of this process (multirank.py), so a handle is the pointer itself
// (SIMT_IPC=0: not supported, as on a machine without peer access — the library then exchanges through NCCL)
static inline bool simt_ipc_on() {
  static const bool on = [] { const char* e = getenv("SIMT_IPC"); return !e || atoi(e) != 0; }();
  return on;
}
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  if (!simt_ipc_on()) return cudaErrorNotSupported;
  memset(h, 0, sizeof *h);
  memcpy(h->reserved, &p, sizeof p);
  return cudaSuccess;
}
static inline cudaError_t cudaIpcOpenMemHandle(void** out, cudaIpcMemHandle_t h, unsigned) {
  if (!simt_ipc_on()) return cudaErrorNotSupported;
  memcpy(out, h.reserved, sizeof *out);
  return cudaSuccess;
}
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

template <typename... KArgs, typename... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args&&... args) {
  std::tuple<KArgs...> t(std::forward<Args>(args)...);
  std::function<void()> body = [&]() { std::apply(kernel, t); };
  simt::run_grid(cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, body);
  return cudaSuccess;
}
