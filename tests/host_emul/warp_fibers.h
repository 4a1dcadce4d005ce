// TEST INFRASTRUCTURE: a 32-lane warp on the CPU, for device code that cooperates through warp collectives
// (genarchbench_b200/csrc/kswv_kernels.cuh). Every lane is a ucontext fiber; a collective publishes the lane's
// value and yields round-robin until every lane of its mask has published, then reads their values. Groups of
// lanes (aligned, power-of-two width) may follow different control flow between full-mask collectives, as they
// may on the GPU; within a mask every lane must reach the same collectives in the same order.
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
    uint32_t slot[2][2][kLanes];     // [scope: 0 = a group's mask, 1 = the full mask][parity][lane]
    uint32_t gen[2][kLanes];
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
        w.gen[0][l] = w.gen[1][l] = 0;
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

// switch to the next lane that has not returned yet (round robin); returns when this lane is scheduled again
inline void yield_lane() {
    Warp *w = current();
    const int me = w->cur;
    int nxt = me;
    for (int step = 1; step <= Warp::kLanes; ++step) {
        const int cand = (me + step) % Warp::kLanes;
        if (!w->done[cand]) { nxt = cand; break; }
    }
    if (nxt == me) return;
    w->cur = nxt;
    swapcontext(&w->ctx[me], &w->ctx[nxt]);
}

// A collective over the lanes in `mask` (which must contain the caller, as on the GPU): publish v, wait until every
// lane of the mask has published its value of the same collective, return the buffer the values sit in. Lanes
// outside the mask are free to be anywhere else in the program (groups of a warp may diverge). Collectives are
// counted per lane in two scopes, a group's mask and the full mask: the lanes of a group run the same code, so their
// group counts agree; every lane reaches the full-mask collectives at the same program points, so those agree too.
inline const uint32_t *publish(uint32_t v, uint32_t mask = 0xFFFFFFFFu) {
    Warp *w = current();
    const int me = w->cur;
    const int scope = mask == 0xFFFFFFFFu ? 1 : 0;
    const uint32_t g = ++w->gen[scope][me];
    uint32_t *buf = w->slot[scope][g & 1u];
    buf[me] = v;
    for (int l = 0; l < Warp::kLanes; ++l) {
        if (!((mask >> l) & 1u) || l == me) continue;
        while (w->gen[scope][l] < g) {
            if (w->done[l]) break;      // a bug in the code under test; do not hang
            yield_lane();
        }
    }
    return buf;
}

inline uint32_t shfl_up1(uint32_t v, uint32_t mask = 0xFFFFFFFFu, int width = 32) {
    const int me = lane();
    const uint32_t *b = publish(v, mask);
    return (me & (width - 1)) > 0 ? b[me - 1] : v;
}
inline uint32_t shfl(uint32_t v, int src, uint32_t mask = 0xFFFFFFFFu, int width = 32) {
    const int me = lane();
    const uint32_t *b = publish(v, mask);
    return b[(me & ~(width - 1)) + (src & (width - 1))];
}
inline uint32_t ballot(bool p, uint32_t mask = 0xFFFFFFFFu) {
    const uint32_t *b = publish(p ? 1u : 0u, mask);
    uint32_t r = 0;
    for (int l = 0; l < Warp::kLanes; ++l) if ((mask >> l) & 1u) r |= (b[l] & 1u) << l;
    return r;
}
inline bool any(bool p, uint32_t mask = 0xFFFFFFFFu) { return ballot(p, mask) != 0; }
inline uint32_t reduce_max(uint32_t v, uint32_t mask = 0xFFFFFFFFu) {
    const uint32_t *b = publish(v, mask);
    uint32_t r = 0;
    for (int l = 0; l < Warp::kLanes; ++l) if (((mask >> l) & 1u) && b[l] > r) r = b[l];
    return r;
}
inline void syncwarp(uint32_t mask = 0xFFFFFFFFu) { publish(0, mask); }

}  // namespace wf
