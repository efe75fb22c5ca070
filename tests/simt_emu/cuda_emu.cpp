// Fiber scheduler of the SIMT-on-host shim (see cuda_emu.h).  Test infrastructure only.
#include "cuda_emu.h"

uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);

namespace svae_emu {

static constexpr size_t kStack = 256 * 1024;

State& state() {
    static State s;
    return s;
}

void* dyn_smem() { return state().dyn; }

unsigned long long globaltimer() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

void yield(int why) {
    State& s = state();
    const int me = s.cur;
    s.status[me] = why;
    swapcontext(&s.ctx[me], &s.sched);
}

static void trampoline() {
    State& s = state();
    (*s.body)();
    s.status[s.cur] = State::DONE;
    swapcontext(&s.ctx[s.cur], &s.sched);
}

void run_block(std::function<void()>& body, int n) {
    State& s = state();
    s.body = &body;
    s.n = n;
    if ((int)s.ctx.size() < n) {
        s.ctx.resize(n);
        s.status.resize(n);
        s.shfl.resize(n);
        while ((int)s.stacks.size() < n) s.stacks.push_back((char*)malloc(kStack));
    }
    for (int i = 0; i < n; ++i) {
        getcontext(&s.ctx[i]);
        s.ctx[i].uc_stack.ss_sp = s.stacks[i];
        s.ctx[i].uc_stack.ss_size = kStack;
        s.ctx[i].uc_link = nullptr;
        makecontext(&s.ctx[i], trampoline, 0);
        s.status[i] = State::RUNNABLE;
    }
    for (;;) {
        bool ran = false;
        for (int i = 0; i < n; ++i) {
            if (s.status[i] != State::RUNNABLE) continue;
            s.cur = i;
            threadIdx = uint3{(unsigned)i, 0, 0};
            swapcontext(&s.sched, &s.ctx[i]);
            ran = true;
        }
        // release the barriers every live thread of their scope has reached
        bool released = false, all_done = true, block_ready = true, any_block = false;
        for (int i = 0; i < n; ++i) {
            if (s.status[i] == State::DONE) continue;
            all_done = false;
            if (s.status[i] == State::WAIT_BLOCK) any_block = true; else block_ready = false;
        }
        if (all_done) return;
        if (any_block && block_ready) {
            for (int i = 0; i < n; ++i) if (s.status[i] == State::WAIT_BLOCK) s.status[i] = State::RUNNABLE;
            released = true;
        }
        for (int w = 0; w * 32 < n; ++w) {
            bool ready = true, any = false;
            for (int i = w * 32; i < n && i < w * 32 + 32; ++i) {
                if (s.status[i] == State::DONE) continue;
                if (s.status[i] == State::WAIT_WARP) any = true; else ready = false;
            }
            if (any && ready) {
                for (int i = w * 32; i < n && i < w * 32 + 32; ++i)
                    if (s.status[i] == State::WAIT_WARP) s.status[i] = State::RUNNABLE;
                released = true;
            }
        }
        if (!ran && !released) {
            fprintf(stderr, "svae_emu: barrier deadlock in block (%u,%u,%u): divergent __syncthreads / shuffle\n",
                    blockIdx.x, blockIdx.y, blockIdx.z);
            abort();
        }
    }
}

}  // namespace svae_emu
