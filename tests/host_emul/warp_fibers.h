// TEST INFRASTRUCTURE: a 32-lane warp on the CPU, for device code that cooperates through warp collectives
// (genarchbench_b200/csrc/kswv_kernels.cuh). Every lane is a ucontext fiber; a collective publishes the lane's
// value, yields round-robin, and reads the other lanes' values once every lane has arrived. The code under test
// must reach the same collectives in the same order on every lane (it does on the GPU too: all of them are
// called with the full mask) and must leave the warp function on all lanes after the same collective.
#pragma once
#include <stdint.h>
#include <ucontext.h>
#include <functional>
#include <vector>

namespace wf {

struct Warp {
    static constexpr int kLanes = 32;
    static constexpr size_t kStack = 256 * 1024;
    ucontext_t main_ctx, ctx[kLanes];
    std::vector<char> stacks;
    int cur = 0;
    bool done[kLanes];
    uint32_t slot[2][kLanes];
    uint32_t gen[kLanes];
    std::function<void()> fn;
    Warp() : stacks(kStack * kLanes) {}
};

inline Warp *&current() {
    static thread_local Warp *w = nullptr;
    return w;
}

inline void fiber_entry() {
    Warp *w = current();
    w->fn();
    w->done[w->cur] = true;
}

// Runs fn on 32 lanes in lock step; returns when every lane has returned.
inline void run_warp(Warp &w, std::function<void()> fn) {
    current() = &w;
    w.fn = std::move(fn);
    for (int l = 0; l < Warp::kLanes; ++l) {
        w.done[l] = false;
        w.gen[l] = 0;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stacks.data() + Warp::kStack * (size_t)l;
        w.ctx[l].uc_stack.ss_size = Warp::kStack;
        w.ctx[l].uc_link = &w.main_ctx;
        makecontext(&w.ctx[l], (void (*)())fiber_entry, 0);
    }
    for (;;) {
        int l = 0;
        while (l < Warp::kLanes && w.done[l]) ++l;
        if (l == Warp::kLanes) break;
        w.cur = l;
        swapcontext(&w.main_ctx, &w.ctx[l]);
    }
    current() = nullptr;
}

inline int lane() { return current()->cur; }

// publish v, let every other lane publish, return the buffer all 32 values sit in
inline const uint32_t *publish(uint32_t v) {
    Warp *w = current();
    const int me = w->cur;
    uint32_t *buf = w->slot[w->gen[me] & 1u];
    buf[me] = v;
    ++w->gen[me];
    const int nxt = (me + 1) % Warp::kLanes;
    w->cur = nxt;
    swapcontext(&w->ctx[me], &w->ctx[nxt]);
    return buf;
}

inline uint32_t shfl_up1(uint32_t v) { const int me = lane(); const uint32_t *b = publish(v); return me > 0 ? b[me - 1] : v; }
inline uint32_t shfl(uint32_t v, int src) { const uint32_t *b = publish(v); return b[src & 31]; }
inline uint32_t ballot(bool p) {
    const uint32_t *b = publish(p ? 1u : 0u);
    uint32_t r = 0;
    for (int l = 0; l < Warp::kLanes; ++l) r |= (b[l] & 1u) << l;
    return r;
}
inline bool any(bool p) { return ballot(p) != 0; }
inline uint32_t reduce_max(uint32_t v) {
    const uint32_t *b = publish(v);
    uint32_t r = 0;
    for (int l = 0; l < Warp::kLanes; ++l) r = b[l] > r ? b[l] : r;
    return r;
}
inline void syncwarp() { publish(0); }

}  // namespace wf
