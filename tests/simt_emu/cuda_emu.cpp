// Fiber scheduler of the SIMT-on-host shim (see cuda_emu.h).  Test infrastructure only.
#include "cuda_emu.h"

#include <algorithm>

uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);

namespace svae_emu {

static constexpr size_t kStack = 256 * 1024;

State& state() {
    static State s;
    return s;
}

void* dyn_smem() { return state().dyn[state().cta()]; }
void* dyn_smem_of(int cta_rank) { return state().dyn[cta_rank]; }

unsigned long long globaltimer() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

static void switch_out(int status) {
    State& s = state();
    const int me = s.cur;
    s.status[me] = status;
    swapcontext(&s.ctx[me], &s.sched);
}

void yield(int why) { switch_out(why); }

void wait_named(int id, int count) {
    State& s = state();
    s.named_id[s.cur] = id;
    s.named_count[s.cur] = count;
    switch_out(WAIT_NAMED);
}

void wait_until(std::function<bool()> cond, const char* why) {
    if (cond()) return;
    State& s = state();
    s.cond[s.cur] = std::move(cond);
    s.why[s.cur] = why;
    switch_out(WAIT_COND);
}

static void trampoline() {
    State& s = state();
    (*s.body)();
    s.status[s.cur] = DONE;
    swapcontext(&s.ctx[s.cur], &s.sched);
}

// all live fibers in [lo, hi) wait with `status` (and, for named barriers, on `id`): release them
static bool release_scope(State& s, int lo, int hi, int status, int id = -1, int need = -1) {
    int waiting = 0, live = 0;
    for (int i = lo; i < hi; ++i) {
        if (s.status[i] == DONE) continue;
        ++live;
        if (s.status[i] == status && (id < 0 || s.named_id[i] == id)) ++waiting;
    }
    if (waiting == 0) return false;
    if (need >= 0 ? waiting < need : waiting < live) return false;
    for (int i = lo; i < hi; ++i)
        if (s.status[i] == status && (id < 0 || s.named_id[i] == id)) s.status[i] = RUNNABLE;
    return true;
}

void run_cluster(std::function<void()>& body, int per_cta, int cluster, unsigned first_block, size_t smem) {
    State& s = state();
    s.body = &body;
    s.per_cta = per_cta;
    s.cluster = cluster;
    s.first_block = first_block;
    s.dyn_req = smem;
    const int n = per_cta * cluster;
    s.n = n;
    if ((int)s.ctx.size() < n) {
        s.ctx.resize(n);
        s.status.resize(n);
        s.named_id.resize(n);
        s.named_count.resize(n);
        s.cond.resize(n);
        s.why.resize(n);
        s.shfl.resize(n);
        while ((int)s.stacks.size() < n) s.stacks.push_back((char*)malloc(kStack));
    }
    if (smem > s.dyn_cap || (int)s.dyn.size() < cluster) {
        for (void* p : s.dyn) free(p);
        s.dyn.clear();
        s.dyn_cap = smem > s.dyn_cap ? smem : s.dyn_cap;
        for (int c = 0; c < (cluster > 2 ? cluster : 2); ++c)
            s.dyn.push_back(aligned_alloc(1024, (s.dyn_cap + 1023) / 1024 * 1024 + 1024));
    }
    // dynamic shared memory starts out as NaN patterns for every block: reads of bytes the block never wrote show up
    for (int c = 0; c < cluster; ++c) memset(s.dyn[c], 0xFF, smem);
    for (int i = 0; i < n; ++i) {
        getcontext(&s.ctx[i]);
        s.ctx[i].uc_stack.ss_sp = s.stacks[i];
        s.ctx[i].uc_stack.ss_size = kStack;
        s.ctx[i].uc_link = nullptr;
        makecontext(&s.ctx[i], trampoline, 0);
        s.status[i] = RUNNABLE;
    }
    // SVAE_EMU_SHUFFLE=<seed>: visit the fibers in a different pseudo-random order on every pass instead of by thread
    // index, to shake out results that depend on which warp happens to run first between two barriers
    static const char* shuffle_env = getenv("SVAE_EMU_SHUFFLE");
    static unsigned long long rng = shuffle_env ? strtoull(shuffle_env, nullptr, 10) * 2654435761ull + 1 : 0;
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    for (;;) {
        bool progress = false;
        if (shuffle_env)
            for (int i = n - 1; i > 0; --i) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(order[i], order[(int)((rng >> 33) % (unsigned long long)(i + 1))]);
            }
        for (int oi = 0; oi < n; ++oi) {
            const int i = order[oi];
            if (s.status[i] == WAIT_COND && s.cond[i]()) s.status[i] = RUNNABLE;
            if (s.status[i] != RUNNABLE) continue;
            s.cur = i;
            threadIdx = uint3{(unsigned)(i % per_cta), 0, 0};
            blockIdx.x = first_block + (unsigned)(i / per_cta);
            swapcontext(&s.sched, &s.ctx[i]);
            progress = true;
        }
        bool all_done = true;
        for (int i = 0; i < n; ++i) all_done = all_done && s.status[i] == DONE;
        if (all_done) return;
        // release the barriers every live thread of their scope has reached
        for (int c = 0; c < cluster; ++c) {
            const int lo = c * per_cta, hi = lo + per_cta;
            progress |= release_scope(s, lo, hi, WAIT_BLOCK);
            for (int w = lo; w < hi; w += 32) progress |= release_scope(s, w, w + 32 < hi ? w + 32 : hi, WAIT_WARP);
            for (int i = lo; i < hi; ++i)
                if (s.status[i] == WAIT_NAMED) progress |= release_scope(s, lo, hi, WAIT_NAMED, s.named_id[i], s.named_count[i]);
        }
        progress |= release_scope(s, 0, n, WAIT_CLUSTER);
        if (!progress) {
            bool cond_ready = false;
            for (int i = 0; i < n; ++i) cond_ready = cond_ready || (s.status[i] == WAIT_COND && s.cond[i]());
            if (cond_ready) continue;
            fprintf(stderr, "svae_emu: deadlock in cluster at block %u:\n", first_block);
            static const char* names[] = {"runnable", "__syncthreads", "warp barrier", "named barrier", "cluster barrier",
                                          "condition", "done"};
            int shown = 0;
            for (int i = 0; i < n && shown < 24; ++i)
                if (s.status[i] != DONE) {
                    fprintf(stderr, "  cta %d thread %d: %s %s\n", i / per_cta, i % per_cta, names[s.status[i]],
                            s.status[i] == WAIT_COND && s.why[i] ? s.why[i] : "");
                    ++shown;
                }
            abort();
        }
    }
}

}  // namespace svae_emu

// ---- cuTensorMapEncodeTiled stand-in (2-D tiled maps, see include/cuda.h) -------------------------------------------
#include "include/cuda.h"
CUresult svae_emu_encode_tiled(CUtensorMap* map, CUtensorMapDataType type, cuuint32_t rank, void* base,
                               const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle swizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    if (rank != 2) return 1;
    memset(map, 0, sizeof(*map));
    map->base = (char*)base;
    map->cols = dims[0];
    map->rows = dims[1];
    map->row_pitch = strides[0];
    map->box_cols = box[0];
    map->box_rows = box[1];
    map->esize = type == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 2;
    map->swizzle = swizzle;
    if (((uintptr_t)base & 15) || (strides[0] & 15) || box[0] * map->esize > 128 || box[1] > 256) return 1;   // driver checks
    return CUDA_SUCCESS;
}
