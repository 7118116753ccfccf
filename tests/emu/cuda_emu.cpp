// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {

static State g_state;
State& st() { return g_state; }

static const size_t kStack = 256 * 1024;

static void set_thread(unsigned t) {
  State& s = g_state;
  s.cur = t;
  threadIdx.x = t % s.block.x;
  threadIdx.y = (t / s.block.x) % s.block.y;
  threadIdx.z = t / (s.block.x * s.block.y);
}

static void release_barriers_on_exit();

static void fiber_entry() {
  State& s = g_state;
  (*s.body)();
  s.done[s.cur] = 1;
  s.alive--;
  release_barriers_on_exit();
  swapcontext(&s.ctx[s.cur], &s.main_ctx);
}

void yield() {
  State& s = g_state;
  unsigned me = s.cur;
  swapcontext(&s.ctx[me], &s.main_ctx);
  set_thread(me);
}

unsigned lanes_in_warp(unsigned warp) {
  State& s = g_state;
  unsigned lo = warp * 32;
  unsigned hi = std::min(lo + 32, s.nthreads);
  return hi - lo;
}

static void release_barriers_on_exit() {
  // exited threads count as arrived (Volta+ semantics)
  State& s = g_state;
  if (s.alive > 0 && s.bar_count >= s.alive) {
    s.bar_count = 0;
    s.bar_gen++;
  }
}

void syncthreads() {
  State& s = g_state;
  unsigned gen = s.bar_gen;
  if (++s.bar_count >= s.alive) {
    s.bar_count = 0;
    s.bar_gen++;
    return;
  }
  while (s.bar_gen == gen) yield();
}

void warp_sync(unsigned warp, unsigned nlanes) {
  State& s = g_state;
  unsigned gen = s.warp_gen[warp];
  if (++s.warp_count[warp] >= nlanes) {
    s.warp_count[warp] = 0;
    s.warp_gen[warp]++;
    return;
  }
  while (s.warp_gen[warp] == gen) yield();
}

uint64_t shfl_generic(uint64_t v, int arg, int mode, int width) {
  State& s = g_state;
  unsigned tid = s.cur;
  unsigned warp = tid / 32, lane = tid % 32;
  unsigned nl = lanes_in_warp(warp);
  unsigned nwarps = (s.nthreads + 31) / 32;
  unsigned parity = s.warp_gen[warp] & 1u;
  uint64_t* slots = &s.warp_slots[(size_t(parity) * nwarps + warp) * 32];
  slots[lane] = v;
  warp_sync(warp, nl);
  int src;
  int base = (lane / width) * width;
  int rel = lane % width;
  switch (mode) {
    case 0: src = base + (arg % width); break;
    case 1: src = base + ((rel ^ arg) % width); if ((rel ^ arg) >= width) src = lane; break;
    case 2: src = (rel + arg < width) ? int(lane) + arg : int(lane); break;
    default: src = (rel - arg >= 0) ? int(lane) - arg : int(lane); break;
  }
  if (src < 0 || unsigned(src) >= nl) src = lane;
  return slots[src];
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  State& s = g_state;
  s.grid = grid;
  s.block = block;
  gridDim = grid;
  blockDim = block;
  s.nthreads = block.x * block.y * block.z;
  unsigned nwarps = (s.nthreads + 31) / 32;
  s.body = &body;
  if (s.stacks.size() < s.nthreads) {
    size_t old = s.stacks.size();
    s.stacks.resize(s.nthreads);
    for (size_t i = old; i < s.nthreads; ++i) s.stacks[i] = static_cast<char*>(std::malloc(kStack));
  }
  static const int order = [] {
    const char* e = std::getenv("LMVN_EMU_ORDER");
    return !e ? 0 : (e[0] == 'r' ? 1 : (e[0] == 's' ? 2 : 0));
  }();
  static std::vector<unsigned> perm;
  if (order == 2 && perm.size() != s.nthreads) {
    perm.resize(s.nthreads);
    for (unsigned i = 0; i < s.nthreads; ++i) perm[i] = i;
    unsigned long long x = 0x2545F4914F6CDD1Dull;  // fixed seed: runs are reproducible
    for (unsigned i = s.nthreads; i > 1; --i) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      std::swap(perm[i - 1], perm[x % i]);
    }
  }
  s.ctx.resize(s.nthreads);
  s.done.assign(s.nthreads, 0);
#ifdef __SANITIZE_ADDRESS__
  // exact, fresh allocation: AddressSanitizer then sees every access past the shared memory a launch asked for
  std::vector<unsigned char>(smem ? smem : 1).swap(s.dyn_smem);
#else
  s.dyn_smem.resize(smem + 16);
#endif
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        std::fill(s.dyn_smem.begin(), s.dyn_smem.end(), 0xff);  // NaN poison
        s.alive = s.nthreads;
        s.bar_count = 0;
        s.warp_count.assign(nwarps, 0);
        s.warp_gen.assign(nwarps, 0);
        s.warp_slots.assign(size_t(2) * nwarps * 32, 0);
        std::fill(s.done.begin(), s.done.end(), 0);
        for (unsigned t = 0; t < s.nthreads; ++t) {
          getcontext(&s.ctx[t]);
          s.ctx[t].uc_stack.ss_sp = s.stacks[t];
          s.ctx[t].uc_stack.ss_size = kStack;
          s.ctx[t].uc_link = &s.main_ctx;
          makecontext(&s.ctx[t], fiber_entry, 0);
        }
        unsigned remaining = s.nthreads;
        unsigned long spins = 0;
        while (remaining) {
          unsigned before = remaining;
          for (unsigned k = 0; k < s.nthreads; ++k) {
            // LMVN_EMU_ORDER=reverse | shuffle: threads run one after the other up to their next barrier, so a missing
            // barrier shows as a stale or poisoned read under one of the orders (the race check of the shared-memory
            // exchanges; tools/run_emu_sanitized.py --order)
            const unsigned t = order == 1 ? s.nthreads - 1 - k : (order == 2 ? perm[k] : k);
            if (s.done[t]) continue;
            set_thread(t);
            swapcontext(&s.main_ctx, &s.ctx[t]);
            if (s.done[t]) remaining--;
          }
          if (remaining == before && ++spins > 100000000ul) {
            std::fprintf(stderr, "[cuda_emu] deadlock suspected in block (%u,%u,%u)\n", bx, by, bz);
            std::abort();
          }
        }
      }
  s.body = nullptr;
}

}  // namespace emu
