// Lock-step emulation of one CUDA warp on the host (test infrastructure, never linked into the product).
//
// The 32 lanes run as ucontext fibres inside one thread.  Every warp-collective intrinsic is a rendezvous:
// a lane publishes its operand, yields until all 32 lanes have arrived, then reads what it needs.  That is
// the semantics of the *_sync intrinsics for converged callers, which is how sparsex_b200/csrc/chunk_kernel.cuh
// uses them (all collectives sit in warp-uniform control flow).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __device__
#define __forceinline__ inline
#define __align__(n) alignas(n)

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
using std::min;
using std::max;
using std::fma;

namespace warp_emul {
constexpr int LANES = 32;
struct Warp {
  ucontext_t sched, ctx[LANES];
  char *stack[LANES];
  bool done[LANES];
  int lane = 0, arrived = 0;
  unsigned phase = 0;
  uint64_t slot[2][LANES];
  std::function<void(int)> body;
};
inline Warp *&cur() { static Warp *w = nullptr; return w; }
inline void yield_() { Warp *w = cur(); swapcontext(&w->ctx[w->lane], &w->sched); }
inline void trampoline() {
  Warp *w = cur();
  w->body(w->lane);
  w->done[w->lane] = true;
  swapcontext(&w->ctx[w->lane], &w->sched);
}
// Runs body(lane) for lanes 0..31 in lock step at the collectives.
inline void run_warp(const std::function<void(int)> &body) {
  static Warp W;
  static bool stacks = false;
  const size_t STK = 256 * 1024;
  if (!stacks) { for (int i = 0; i < LANES; i++) W.stack[i] = (char *)malloc(STK); stacks = true; }
  W.body = body; W.arrived = 0; W.phase = 0;
  cur() = &W;
  for (int i = 0; i < LANES; i++) {
    W.done[i] = false;
    getcontext(&W.ctx[i]);
    W.ctx[i].uc_stack.ss_sp = W.stack[i];
    W.ctx[i].uc_stack.ss_size = STK;
    W.ctx[i].uc_link = &W.sched;
    makecontext(&W.ctx[i], (void (*)())trampoline, 0);
  }
  for (;;) {
    bool any = false;
    for (int i = 0; i < LANES; i++) {
      if (W.done[i]) continue;
      any = true;
      W.lane = i;
      swapcontext(&W.sched, &W.ctx[i]);
    }
    if (!any) break;
  }
}
template <class T>
inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
// publish `bits`, wait for the whole warp, return the 32 published operands
inline const uint64_t *exchange(uint64_t bits) {
  Warp *w = cur();
  const unsigned ph = w->phase;
  const int me = w->lane;
  w->slot[ph & 1][me] = bits;
  if (++w->arrived == LANES) { w->arrived = 0; w->phase++; }
  else while (w->phase == ph) { yield_(); }
  w->lane = me;
  return w->slot[ph & 1];
}
inline int lane_id() { return cur()->lane; }
// threadIdx of the emulated thread: the harness sets the warp's position in its CTA before run_warp()
inline int &warp_in_cta() { static int w = 0; return w; }
struct Tid { int x; };
inline Tid tid() { return Tid{warp_in_cta() * 32 + lane_id()}; }
}  // namespace warp_emul
#define threadIdx (warp_emul::tid())

template <class T>
inline T __shfl_sync(unsigned, T v, int src) { const uint64_t *s = warp_emul::exchange(warp_emul::to_bits(v)); return warp_emul::from_bits<T>(s[src & 31]); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, int o) {
  const int me = warp_emul::lane_id();
  const uint64_t *s = warp_emul::exchange(warp_emul::to_bits(v));
  return warp_emul::from_bits<T>(s[me - o >= 0 ? me - o : me]);
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, int o) {
  const int me = warp_emul::lane_id();
  const uint64_t *s = warp_emul::exchange(warp_emul::to_bits(v));
  return warp_emul::from_bits<T>(s[me + o < 32 ? me + o : me]);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int o) {
  const int me = warp_emul::lane_id();
  const uint64_t *s = warp_emul::exchange(warp_emul::to_bits(v));
  return warp_emul::from_bits<T>(s[(me ^ o) & 31]);
}
inline unsigned __ballot_sync(unsigned, bool p) {
  const uint64_t *s = warp_emul::exchange(p ? 1 : 0);
  unsigned m = 0;
  for (int i = 0; i < 32; i++) m |= (unsigned)(s[i] & 1) << i;
  return m;
}
inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  const uint64_t *s = warp_emul::exchange(v);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r = std::max(r, (unsigned)s[i]);
  return r;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  const uint64_t *s = warp_emul::exchange(v);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)s[i];
  return r;
}
inline unsigned __match_any_sync(unsigned, unsigned v) {
  const uint64_t *s = warp_emul::exchange(v);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)((unsigned)s[i] == v) << i;
  return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline void __syncwarp() { warp_emul::exchange(0); }
template <class T>
inline T __ldg(const T *p) { return *p; }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
  return c;
}
inline double atomicAdd(double *p, double v) { double o = *p; *p = o + v; return o; }
