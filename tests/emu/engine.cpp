// TEST INFRASTRUCTURE ONLY -- fiber scheduler of the warp-lockstep CUDA emulator (see include/cuda_runtime.h).
//
// One OS thread runs one thread block at a time; every CUDA thread of the block is a fiber with its own stack.
// A fiber runs until it reaches a warp collective or __syncthreads, where it parks until all live lanes of its
// warp (threads of its block) named by the mask have arrived.  What real hardware leaves undefined is an ERROR here:
//   * lanes of one warp waiting in DIFFERENT collectives (divergent *_sync calls),
//   * a collective whose mask names a lane that already exited the kernel,
//   * a shuffle that reads a lane which did not take part,
//   * a block in which nobody can make progress (deadlock).
// The error is sticky and surfaces through cudaGetLastError(), i.e. through the library's own LGB_LAUNCH_CHECK.
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include <sys/mman.h>

#include "cuda_runtime.h"

#if !defined(__x86_64__)
#error "the emulator's context switch is written for x86-64"
#endif

extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
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
.size emu_switch,.-emu_switch
)");

namespace emu {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {

constexpr size_t STACK_BYTES = 256 * 1024;   // + one guard page below: an overflow faults instead of corrupting a neighbour
constexpr int MAX_THREADS = 1024;

enum Wait { RUNNABLE = 0, WAIT_WARP = 1, WAIT_BLOCK = 2, DONE = 3 };

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  int tid = 0;
  uint3 tidx{0, 0, 0};
  int wait = DONE;
  unsigned wait_gen = 0;
};

struct Warp {
  uint64_t slot[32];
  uint64_t result[32];
  unsigned arrived_mask = 0, result_mask = 0, alive_mask = 0, want_mask = 0;
  int op = 0;
  unsigned gen = 0;
};

std::mutex g_mu;               // one launch at a time (the engine state is global)
std::vector<Fiber> g_fibers;
Warp g_warps[MAX_THREADS / 32];
void* g_sched_sp = nullptr;
Fiber* g_cur = nullptr;
body_fn g_body = nullptr;
void* g_body_arg = nullptr;
int g_block_alive = 0, g_block_arrived = 0;
unsigned g_block_gen = 0;
std::string g_error;           // sticky
std::string g_last_error_text;
bool g_abort = false;
long g_inactive_reads = 0;

void fail(const std::string& msg) {
  if (g_error.empty()) {
    char where[160];
    snprintf(where, sizeof(where), " [block (%u,%u,%u) thread %d]", g_blockIdx.x, g_blockIdx.y, g_blockIdx.z,
             g_cur ? g_cur->tid : -1);
    g_error = "cuda-emu: " + msg + where;
  }
  g_abort = true;
}

void to_scheduler() { emu_switch(&g_cur->sp, g_sched_sp); }

void complete_warp(Warp& W) {
  memcpy(W.result, W.slot, sizeof(W.slot));
  W.result_mask = W.arrived_mask;
  W.arrived_mask = 0;
  W.want_mask = 0;
  W.op = 0;
  W.gen++;
}

void fiber_exit() {
  Fiber* f = g_cur;
  Warp& W = g_warps[f->tid >> 5];
  const unsigned bit = 1u << (f->tid & 31);
  W.alive_mask &= ~bit;
  if (W.arrived_mask) {
    if (W.want_mask & bit)
      fail("a lane exited the kernel while the rest of its warp waits for it in a *_sync collective (mask names an exited lane)");
    else if ((W.arrived_mask & (W.want_mask & W.alive_mask)) == (W.want_mask & W.alive_mask))
      complete_warp(W);
  }
  g_block_alive--;
  if (g_block_arrived > 0 && g_block_arrived == g_block_alive)
    fail("a thread exited the kernel while the rest of its block waits in __syncthreads()");
  f->wait = DONE;
  to_scheduler();
  abort();   // never resumed
}

void trampoline() {
  g_body(g_body_arg);
  fiber_exit();
}

void prepare(Fiber& f) {
  if (!f.stack) {
    const size_t guard = 4096;
    void* m = mmap(nullptr, STACK_BYTES + guard, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) { perror("cuda-emu: mmap"); abort(); }
    mprotect(m, guard, PROT_NONE);
    f.stack = (char*)m + guard;
  }
  uint64_t* sp = (uint64_t*)(((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15);
  *--sp = 0;                        // return address slot of trampoline()'s imaginary caller (keeps rsp % 16 == 8 at entry)
  *--sp = (uint64_t)&trampoline;    // popped by emu_switch's ret
  for (int i = 0; i < 6; ++i) *--sp = 0;   // rbp rbx r12 r13 r14 r15
  f.sp = sp;
  f.wait = RUNNABLE;
}

uint64_t g_rng = 0x9e3779b97f4a7c15ull;
int order_mode_and_seed() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("LGB_EMU_ORDER");
    mode = (!e || !strncmp(e, "forward", 7)) ? 0 : (!strncmp(e, "reverse", 7) ? 1 : 2);
    if (mode == 2 && e && strchr(e, ':')) g_rng ^= strtoull(strchr(e, ':') + 1, nullptr, 10) * 0xbf58476d1ce4e5b9ull;
  }
  return mode;
}

void run_block(int nthreads) {
  const int nwarps = (nthreads + 31) / 32;
  for (int w = 0; w < nwarps; ++w) {
    Warp& W = g_warps[w];
    W.arrived_mask = W.result_mask = W.want_mask = 0;
    W.op = 0;
    W.gen = 0;
    const int lanes = std::min(32, nthreads - 32 * w);
    W.alive_mask = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1);
  }
  g_block_alive = nthreads;
  g_block_arrived = 0;
  g_block_gen = 0;
  for (int t = 0; t < nthreads; ++t) {
    Fiber& f = g_fibers[t];
    f.tid = t;
    f.tidx.x = t % g_blockDim.x;
    f.tidx.y = (t / g_blockDim.x) % g_blockDim.y;
    f.tidx.z = t / (g_blockDim.x * g_blockDim.y);
    prepare(f);
  }
  // Visiting order of the runnable threads (LGB_EMU_ORDER = forward | reverse | shuffle[:seed]).  Correct kernels cannot
  // tell the difference: anything that communicates through shared or global memory without the barrier / collective it
  // needs computes something else under another order -- a cheap race fuzzer for code the lockstep order would let pass.
  const int order_mode = order_mode_and_seed();
  uint64_t& rng = g_rng;
  static std::vector<int> order;
  order.resize(nthreads);
  for (int t = 0; t < nthreads; ++t) order[t] = order_mode == 1 ? nthreads - 1 - t : t;
  while (g_block_alive > 0 && !g_abort) {
    bool ran = false;
    if (order_mode == 2)
      for (int t = nthreads - 1; t > 0; --t) {          // Fisher-Yates with an xorshift generator, new order every pass
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        std::swap(order[t], order[(int)(rng % (uint64_t)(t + 1))]);
      }
    for (int oi = 0; oi < nthreads && !g_abort; ++oi) {
      const int t = order[oi];
      Fiber& f = g_fibers[t];
      if (f.wait == DONE) continue;
      if (f.wait == WAIT_WARP && g_warps[t >> 5].gen == f.wait_gen) continue;
      if (f.wait == WAIT_BLOCK && g_block_gen == f.wait_gen) continue;
      f.wait = RUNNABLE;
      g_cur = &f;
      g_threadIdx = f.tidx;
      emu_switch(&g_sched_sp, f.sp);
      ran = true;
    }
    if (!ran && g_block_alive > 0 && !g_abort) {
      g_cur = nullptr;
      fail("deadlock: every live thread of the block waits in a collective that can never complete "
           "(divergent __syncthreads / *_sync under a mask that names lanes on another path)");
    }
  }
  g_cur = nullptr;
}

}  // namespace

void launch(dim3 grid, dim3 block, body_fn fn, void* arg) {
  std::lock_guard<std::mutex> lk(g_mu);
  const uint64_t nthreads = (uint64_t)block.x * block.y * block.z;
  if (nthreads == 0 || nthreads > MAX_THREADS || grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 ||
      grid.z > 65535) {
    if (g_error.empty()) g_error = "cuda-emu: invalid launch configuration";
    return;
  }
  if (!g_error.empty()) return;   // sticky error: later launches do not run (like a poisoned CUDA context)
  if (g_fibers.size() < nthreads) g_fibers.resize(nthreads);
  g_body = fn;
  g_body_arg = arg;
  g_blockDim = block;
  g_gridDim = grid;
  g_abort = false;
  // blocks of a grid run in LGB_EMU_ORDER too (forward / reverse / a random cyclic stride): hardware promises no order
  const int mode = order_mode_and_seed();
  uint64_t stride = 1, offset = 0;
  if (mode == 2 && grid.x > 2) {
    g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;
    offset = g_rng % grid.x;
    stride = 1 + (g_rng >> 20) % (grid.x - 1);
    while (std::__gcd<uint64_t>(stride, grid.x) != 1) ++stride;      // coprime stride = a permutation of 0..grid.x-1
  }
  for (unsigned bz = 0; bz < grid.z && !g_abort; ++bz)
    for (unsigned by = 0; by < grid.y && !g_abort; ++by)
      for (unsigned i = 0; i < grid.x && !g_abort; ++i) {
        const unsigned bx = mode == 0 ? i : (mode == 1 ? grid.x - 1 - i : (unsigned)((offset + (uint64_t)i * stride) % grid.x));
        g_blockIdx.x = bx; g_blockIdx.y = by; g_blockIdx.z = bz;
        run_block((int)nthreads);
      }
}

uint64_t warp_collect(int op, unsigned mask, uint64_t v, uint64_t* all, unsigned* arrived_mask) {
  Fiber* f = g_cur;
  Warp& W = g_warps[f->tid >> 5];
  const int lane = f->tid & 31;
  if (!((mask >> lane) & 1u)) fail("a lane called a *_sync collective with a mask that does not name itself");
  if (W.arrived_mask == 0) {
    W.op = op;
    W.want_mask = mask;
    if (mask & ~W.alive_mask & (W.alive_mask | ~0u)) {
      // the mask names lanes that do not exist in this (partial) warp or already exited
      const int nthreads = (int)(g_blockDim.x * g_blockDim.y * g_blockDim.z);
      const int lanes = std::min(32, nthreads - 32 * (f->tid >> 5));
      const unsigned exist = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1);
      if (mask & exist & ~W.alive_mask)
        fail("a *_sync collective names a lane that already exited the kernel (undefined behaviour on hardware)");
    }
  } else if (W.op != op || W.want_mask != mask) {
    fail("lanes of one warp wait in different *_sync collectives (divergent collective calls)");
  }
  if (g_abort) { to_scheduler(); abort(); }
  W.slot[lane] = v;
  W.arrived_mask |= 1u << lane;
  const unsigned need = W.want_mask & W.alive_mask;
  if ((W.arrived_mask & need) == need) {
    complete_warp(W);
  } else {
    f->wait = WAIT_WARP;
    f->wait_gen = W.gen;
    to_scheduler();
  }
  memcpy(all, W.result, sizeof(W.result));
  *arrived_mask = W.result_mask;
  return W.result[lane];
}

void block_barrier() {
  Fiber* f = g_cur;
  g_block_arrived++;
  if (g_block_arrived == g_block_alive) {
    g_block_arrived = 0;
    g_block_gen++;
    return;
  }
  f->wait = WAIT_BLOCK;
  f->wait_gen = g_block_gen;
  to_scheduler();
}

void note_inactive_read(const char* what) {
  g_inactive_reads++;
  fail(std::string(what) + " reads a lane that did not take part in the collective (undefined value on hardware)");
  to_scheduler();
  abort();
}

// ---- emulated NVSwitch multicast object: [mc_base, mc_base + bytes) fans out to `world` peer buffers -------------------
namespace {
struct Multicast { char* base; size_t bytes; std::vector<char*> peers; };
std::vector<Multicast> g_mc;
const Multicast* find_mc(const void* p) {
  for (const Multicast& m : g_mc)
    if ((const char*)p >= m.base && (const char*)p < m.base + m.bytes) return &m;
  return nullptr;
}
}  // namespace

float4 mc_ld_reduce(const float4* mc) {
  const Multicast* m = find_mc(mc);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!m) { fail("multimem.ld_reduce on an address that is not a bound multicast object"); return acc; }
  const size_t off = (const char*)mc - m->base;
  for (char* peer : m->peers) {
    const float4 v = *reinterpret_cast<const float4*>(peer + off);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  return acc;
}

void mc_st(float4* mc, const float4& v) {
  const Multicast* m = find_mc(mc);
  if (!m) { fail("multimem.st on an address that is not a bound multicast object"); return; }
  const size_t off = (const char*)mc - m->base;
  for (char* peer : m->peers) *reinterpret_cast<float4*>(peer + off) = v;
}

cudaError_t take_error() {
  if (g_error.empty()) return cudaSuccess;
  g_last_error_text = g_error;
  g_error.clear();
  return cudaErrorLaunchFailure;
}

const char* error_string() { return g_last_error_text.c_str(); }

}  // namespace emu

extern "C" void emu_multicast_clear() { emu::g_mc.clear(); }

extern "C" void emu_multicast_bind(void* mc_base, size_t bytes, int world, void** peer_bases) {
  emu::Multicast m;
  m.base = (char*)mc_base;
  m.bytes = bytes;
  for (int r = 0; r < world; ++r) m.peers.push_back((char*)peer_bases[r]);
  // a registration is only as alive as the tensors behind it: drop every older one whose key range overlaps the new key
  // (the allocator reuses addresses -- a stale range that still matched would redirect loads/stores into freed peers)
  auto& v = emu::g_mc;
  v.erase(std::remove_if(v.begin(), v.end(), [&](const emu::Multicast& o) { return o.base < m.base + m.bytes && m.base < o.base + o.bytes; }),
          v.end());
  v.push_back(m);
}
