// simt_runtime.cpp — the fiber scheduler behind the SIMT shim (see cuda_runtime.h in this directory).  Development aid.
#include <pthread.h>
#include <stdio.h>
#include <sys/mman.h>

#include <atomic>
#include <thread>
#include <algorithm>
#include <vector>

#include "cuda_runtime.h"

// cooperative context switch (x86-64 System V): save the callee-saved registers on the current stack, store the stack
// pointer, load the other stack pointer, restore, return.
extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");

thread_local uint3 threadIdx, blockIdx, blockDim, gridDim;

namespace simt {

thread_local Block* B = nullptr;
thread_local Fiber* cur = nullptr;

namespace {
constexpr size_t STACK_BYTES = 256 << 10;
constexpr int MAX_THREADS = 1024;
constexpr size_t DYN_SMEM = 232448;

struct Worker {  // per OS thread: fiber stacks and dynamic shared memory, allocated once
  char* stacks = nullptr;
  unsigned char* smem = nullptr;
  Fiber fibers[MAX_THREADS];
  Warp warps[MAX_THREADS / 32];
  Worker() {
    stacks = (char*)mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    smem = (unsigned char*)aligned_alloc(128, DYN_SMEM);
    if (stacks == MAP_FAILED || !smem) {
      fprintf(stderr, "simt: cannot allocate fiber stacks\n");
      abort();
    }
  }
  ~Worker() {
    munmap(stacks, STACK_BYTES * MAX_THREADS);
    free(smem);
  }
};
thread_local Worker* W = nullptr;

void fiber_entry() {
  (*B->body)();
  Fiber* f = cur;
  f->done = true;
  B->nlive--;
  B->warps[f->warp].live &= ~(1u << f->lane);
  for (;;) simt_switch(&f->sp, B->sched_sp);  // never resumed once done
}

void run_block(Block& blk) {
  B = &blk;
  blockIdx = blk.bid;
  blockDim = uint3{blk.bdim.x, blk.bdim.y, blk.bdim.z};
  gridDim = uint3{blk.gdim.x, blk.gdim.y, blk.gdim.z};
  const int nt = blk.nthreads;
  for (int w = 0; w < (nt + 31) / 32; ++w) {
    Warp& wp = blk.warps[w];
    const int lanes = std::min(32, nt - 32 * w);
    wp.live = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
    wp.arrived = 0;
    wp.gen = 0;
  }
  for (int t = 0; t < nt; ++t) {
    Fiber& f = blk.fibers[t];
    f.tid.x = (unsigned)t % blk.bdim.x;
    f.tid.y = ((unsigned)t / blk.bdim.x) % blk.bdim.y;
    f.tid.z = (unsigned)t / (blk.bdim.x * blk.bdim.y);
    f.lane = t & 31;
    f.warp = t >> 5;
    f.done = false;
    // initial frame: six callee-saved registers, then the entry point as return address, then a fake caller slot
    uintptr_t top = (uintptr_t)(W->stacks + STACK_BYTES * (size_t)(t + 1));
    top &= ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // fake return address of fiber_entry (never used)
    *--sp = (void*)&fiber_entry;     // `ret` of simt_switch jumps here
    for (int r = 0; r < 6; ++r) *--sp = nullptr;
    f.sp = sp;
  }
  blk.nlive = nt;
  blk.bar_arrived = 0;
  blk.bar_gen = 0;
  unsigned long long rounds = 0;
  while (blk.nlive > 0) {
    for (int t = 0; t < nt; ++t) {
      Fiber& f = blk.fibers[t];
      if (f.done) continue;
      cur = &f;
      threadIdx = f.tid;
      simt_switch(&blk.sched_sp, f.sp);
    }
    if (++rounds > (1ull << 30)) {
      fprintf(stderr, "simt: CTA (%u,%u,%u) makes no progress (dead-locked barrier or collective)\n", blk.bid.x, blk.bid.y, blk.bid.z);
      abort();
    }
  }
  cur = nullptr;
  B = nullptr;
}
}  // namespace

void yield() { simt_switch(&cur->sp, B->sched_sp); }

unsigned char* dyn_smem() { return W->smem; }

const uint64_t* warp_exchange(uint32_t mask, uint64_t v) {
  Fiber* f = cur;
  Warp& wp = B->warps[f->warp];
  const uint32_t bit = 1u << f->lane;
  const uint32_t g = wp.gen;
  const int p = g & 1;
  wp.vals[p][f->lane] = v;
  wp.arrived |= bit;
  for (;;) {
    if (wp.gen != g) break;  // completed by another lane
    const uint32_t need = mask & wp.live;
    if ((wp.arrived & need) == need) {
      wp.arrived &= ~need;
      wp.gen = g + 1;
      break;
    }
    yield();
  }
  return wp.vals[p];
}

void block_barrier() {
  Block& b = *B;
  const uint32_t g = b.bar_gen;
  b.bar_arrived++;
  for (;;) {
    if (b.bar_gen != g) break;
    if ((int)b.bar_arrived >= b.nlive) {
      b.bar_arrived = 0;
      b.bar_gen = g + 1;
      break;
    }
    yield();
  }
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
  const int nt = (int)(block.x * block.y * block.z);
  if (nt <= 0 || nt > MAX_THREADS || smem > DYN_SMEM) {
    fprintf(stderr, "simt: bad launch configuration (%d threads, %zu bytes of dynamic shared memory)\n", nt, smem);
    abort();
  }
  static const int nworkers = [] {
    const char* e = getenv("SIMT_THREADS");
    int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return n < 1 ? 1 : n;
  }();
  std::atomic<unsigned long long> next{0};
  const int n = (int)std::min<unsigned long long>((unsigned long long)nworkers, nblocks);
  // OpenMP keeps its threads between launches, so the per-thread fiber stacks are set up once
#pragma omp parallel num_threads(n)
  {
    if (!W) W = new Worker();
    Worker& worker = *W;
    for (;;) {
      const unsigned long long i = next.fetch_add(1);  // CTAs start in index order
      if (i >= nblocks) break;
      Block blk;
      blk.bid.x = (unsigned)(i % grid.x);
      blk.bid.y = (unsigned)((i / grid.x) % grid.y);
      blk.bid.z = (unsigned)(i / ((unsigned long long)grid.x * grid.y));
      blk.bdim = block;
      blk.gdim = grid;
      blk.nthreads = nt;
      blk.fibers = worker.fibers;
      blk.warps = worker.warps;
      blk.dyn_smem = worker.smem;
      blk.body = &body;
      run_block(blk);
    }
  }
}

}  // namespace simt
