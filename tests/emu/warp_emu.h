// Host emulation of one CUDA warp as 32 cooperatively scheduled fibers (ucontext).
// TEST INFRASTRUCTURE ONLY: lets the CPU test-suite run the warp-synchronous kernel
// bodies in mpc_blaster_b200/csrc/*.cuh without a GPU.  Never linked into the product.
//
// Model: lanes run round-robin; every warp-collective (sync / shfl) is a yield point.
// Because the kernels are warp-uniform in their sequence of collectives, "lane i has
// reached collective #n" holds for all i when lane 0 resumes after yielding at #n.
#pragma once
#include <ucontext.h>
#include <cstdlib>
#include <functional>
#include <vector>

namespace emu {

struct Warp {
    ucontext_t main_ctx;
    ucontext_t ctx[32];
    std::vector<char> stacks;
    double slot[2][32];
    int parity = 0;
    int cur = 0;
    int alive = 0;
    bool done[32];
    std::function<void()> body;
};

inline Warp *&current()
{
    static thread_local Warp *w = nullptr;
    return w;
}

inline int lane() { return current()->cur; }

inline void yield_next()
{
    Warp *w = current();
    int me = w->cur;
    int nxt = me;
    for (int i = 1; i <= 32; i++) {
        int c = (me + i) & 31;
        if (!w->done[c]) { nxt = c; break; }
    }
    if (nxt == me) return;
    w->cur = nxt;
    swapcontext(&w->ctx[me], &w->ctx[nxt]);
}

inline void sync() { yield_next(); }

inline double shfl(double v, int src)
{
    Warp *w = current();
    // the parity flips once per collective, on lane 0's arrival (first to arrive)
    if (w->cur == 0) w->parity ^= 1;
    int p = w->parity;
    w->slot[p][w->cur] = v;
    yield_next();
    // every lane has written slot[p]; the next collective writes slot[p^1]
    return w->slot[p][src & 31];
}

inline void trampoline()
{
    Warp *w = current();
    w->body();
    int me = w->cur;
    w->done[me] = true;
    w->alive--;
    if (w->alive == 0) {
        setcontext(&w->main_ctx);
    }
    // hand over to the next live lane, never to return
    for (int i = 1; i <= 32; i++) {
        int c = (me + i) & 31;
        if (!w->done[c]) { w->cur = c; setcontext(&w->ctx[c]); }
    }
}

// Run `body` once per lane (body reads emu::lane()).  All lanes must execute the same
// sequence of collectives.
inline void run_warp(const std::function<void()> &body, size_t stack_bytes = 1 << 20)
{
    Warp w;
    w.body = body;
    w.stacks.resize(32 * stack_bytes);
    w.alive = 32;
    for (int i = 0; i < 32; i++) {
        w.done[i] = false;
        getcontext(&w.ctx[i]);
        w.ctx[i].uc_stack.ss_sp = w.stacks.data() + (size_t)i * stack_bytes;
        w.ctx[i].uc_stack.ss_size = stack_bytes;
        w.ctx[i].uc_link = nullptr;
        makecontext(&w.ctx[i], (void (*)())trampoline, 0);
    }
    Warp *prev = current();
    current() = &w;
    w.cur = 0;
    swapcontext(&w.main_ctx, &w.ctx[0]);
    current() = prev;
}

}  // namespace emu
