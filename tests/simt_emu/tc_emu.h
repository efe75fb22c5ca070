// Host model of the Blackwell pieces tc_gemm.cu drives through inline PTX: mbarriers, TMA tiled loads / stores with
// the 128-byte swizzle, tensor memory, tcgen05.mma (single CTA and CTA pair) with shared-memory matrix descriptors,
// tcgen05.commit / ld, named barriers and the cluster primitives.  tests/simt_emu/build.py puts this file in place of
// the "PTX wrappers" section of tc_gemm.cu: the function names and signatures are the kernel's own.
//
// What validates the MODEL: the kernel was debugged on a real B200, so its descriptors, swizzled addresses, barrier
// counts and phase arithmetic are known to be right; if this model disagreed with the hardware on any of them, the
// unmodified kernel would compute wrong results (or deadlock) here.  Asynchrony is collapsed: a TMA copy or an MMA
// completes at issue, which is one of the orders the hardware allows.  Test infrastructure only.
#pragma once
#include <map>

namespace svae {
namespace {

// ---- addresses: shared::cluster window = (cta rank << 24) | byte offset in that CTA's dynamic shared memory ----
inline uint32_t smem_u32(const void* p) {
    svae_emu::State& s = svae_emu::state();
    const int c = s.cta();
    const ptrdiff_t off = (const char*)p - (const char*)svae_emu::dyn_smem_of(c);
    if (off < 0 || off >= (ptrdiff_t)(1 << 18)) { fprintf(stderr, "tc_emu: pointer outside the CTA's shared memory\n"); abort(); }
    return ((uint32_t)c << 24) | (uint32_t)off;
}
inline char* smem_ptr(uint32_t addr) {
    if ((addr & 0xFFFFFF) >= svae_emu::state().dyn_req || (int)(addr >> 24) >= svae_emu::state().cluster) {
        fprintf(stderr, "tc_emu: shared-memory address %#x outside the %zu bytes the launch asked for\n", addr,
                svae_emu::state().dyn_req);
        abort();
    }
    return (char*)svae_emu::dyn_smem_of((int)(addr >> 24)) + (addr & 0xFFFFFF);
}
inline uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }      // SWIZZLE_128B on address bits

// ---- mbarrier -------------------------------------------------------------------------------------------------------
struct MBar { int count = 0, pending = 0; long tx = 0; uint32_t phase = 0; };
inline std::map<uint32_t, MBar>& mbars() { static std::map<uint32_t, MBar> m; return m; }
inline void mbar_check(MBar& b) {
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; }
}
inline MBar& mbar_at(uint32_t bar) {
    auto it = mbars().find(bar);
    if (it == mbars().end()) { fprintf(stderr, "tc_emu: mbarrier %#x used before mbarrier.init\n", bar); abort(); }
    return it->second;
}
inline void mbar_init(uint32_t bar, uint32_t count) { MBar b; b.count = b.pending = (int)count; mbars()[bar] = b; }
inline void mbar_arrive(uint32_t bar) { MBar& b = mbar_at(bar); --b.pending; mbar_check(b); }
inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) { MBar& b = mbar_at(bar); b.tx += bytes; --b.pending; mbar_check(b); }
inline void mbar_complete_tx(uint32_t bar, uint32_t bytes) { MBar& b = mbar_at(bar); b.tx -= bytes; mbar_check(b); }
inline void mbar_wait(uint32_t bar, uint32_t parity) {
    mbar_at(bar);
    svae_emu::wait_until([bar, parity]() { return mbar_at(bar).phase != parity; }, "mbarrier.try_wait");
}
inline void fence_barrier_init() {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void fence_proxy_async() {}
inline void tma_prefetch_desc(const CUtensorMap*) {}

// ---- TMA ------------------------------------------------------------------------------------------------------------
// box rows are 128 bytes wide in every map the library builds; row r of the box lands at dst + 128 r, swizzled
inline void tma_copy_in(uint32_t dst, const CUtensorMap* m, int c0, int c1) {
    const uint32_t row_bytes = m->box_cols * m->esize;
    if (row_bytes != 128 || m->swizzle != CU_TENSOR_MAP_SWIZZLE_128B) { fprintf(stderr, "tc_emu: unsupported tensor map\n"); abort(); }
    for (uint32_t r = 0; r < m->box_rows; ++r)
        for (uint32_t e = 0; e < m->box_cols; ++e) {
            const long row = (long)c1 + r, col = (long)c0 + e;
            const uint32_t a = (dst & 0xFFFFFF) + r * 128 + e * m->esize;
            char* d = smem_ptr((dst & 0xFF000000u) | swz128(a));
            if (row >= 0 && col >= 0 && row < (long)m->rows && col < (long)m->cols)
                memcpy(d, m->base + row * m->row_pitch + col * m->esize, m->esize);
            else
                memset(d, 0, m->esize);                       // out-of-bounds elements read as zero
        }
}
inline void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    tma_copy_in(dst, map, c0, c1);
    mbar_complete_tx(bar, map->box_rows * map->box_cols * map->esize);
}
inline void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    if ((dst & 15) || (bytes & 15) || ((uintptr_t)src & 15)) { fprintf(stderr, "tc_emu: misaligned bulk copy\n"); abort(); }
    memcpy(smem_ptr(dst), src, bytes);
    mbar_complete_tx(bar, bytes);
}
inline void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    tma_load_2d(dst, map, leader_bar, c0, c1);
}
// TMA stores are modelled as LATE as the program allows: the copy out of shared memory is only performed when the
// issuing thread's cp.async.bulk.wait_group.read requires it (or at wait_group 0).  A staging buffer that is
// overwritten before that point therefore corrupts the output here, as it may on the hardware.
struct PendingStore { CUtensorMap map; uint32_t src; int c0, c1; };
struct StoreQueue { std::vector<std::vector<PendingStore>> groups; std::vector<PendingStore> open; };
inline StoreQueue& store_queue() {
    static std::map<int, StoreQueue> q;
    return q[svae_emu::state().cur];
}
inline void tma_store_perform(const PendingStore& p) {
    const CUtensorMap* m = &p.map;
    for (uint32_t r = 0; r < m->box_rows; ++r)
        for (uint32_t e = 0; e < m->box_cols; ++e) {
            const long row = (long)p.c1 + r, col = (long)p.c0 + e;
            if (row >= (long)m->rows || col >= (long)m->cols) continue;        // clipped
            const uint32_t a = (p.src & 0xFFFFFF) + r * 128 + e * m->esize;
            memcpy(m->base + row * m->row_pitch + col * m->esize, smem_ptr((p.src & 0xFF000000u) | swz128(a)), m->esize);
        }
}
inline void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    store_queue().open.push_back(PendingStore{*m, src, c0, c1});
}
inline void tma_store_commit() {
    StoreQueue& q = store_queue();
    q.groups.push_back(std::move(q.open));
    q.open.clear();
}
template <int N> inline void tma_store_wait_read() {          // at most N of the most recent groups may still be reading
    StoreQueue& q = store_queue();
    while ((int)q.groups.size() > N) {
        for (const PendingStore& p : q.groups.front()) tma_store_perform(p);
        q.groups.erase(q.groups.begin());
    }
}

// ---- cluster ----------------------------------------------------------------------------------------------------------
inline uint32_t cluster_ctarank() { return (uint32_t)svae_emu::state().cta(); }
inline void cluster_sync_all() { svae_emu::yield(svae_emu::WAIT_CLUSTER); }
inline uint32_t map_to_cta(uint32_t addr, uint32_t rank) { return (addr & 0xFFFFFF) | (rank << 24); }
inline void mbar_arrive_cluster(uint32_t cluster_addr) { mbar_arrive(cluster_addr); }
inline void epi_bar_sync(int group) { svae_emu::wait_named(1 + group, 128); }
inline void epi_bar_sync_all(int nthreads) { svae_emu::wait_named(3, nthreads); }

// ---- tensor memory: 128 lanes x 512 fp32 columns per CTA ----------------------------------------------------------------
inline float* tmem_of(int cta) {
    static std::vector<float> t[2];
    if (t[cta].empty()) t[cta].assign(128 * 512, 0.f);
    return t[cta].data();
}
inline void tmem_alloc(uint32_t dst_smem, uint32_t) { *(uint32_t*)smem_ptr(dst_smem) = 0u; }
inline void tmem_alloc_pair(uint32_t dst_smem, uint32_t) {              // both CTAs execute it; each writes its own slot
    *(uint32_t*)smem_ptr(dst_smem) = 0u;
}
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tmem_dealloc_pair(uint32_t, uint32_t) {}
inline void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    svae_emu::State& s = svae_emu::state();
    const int lane = (int)(taddr >> 16) + (s.tid() & 31);
    const int col = (int)(taddr & 0xFFFF);
    if ((int)(taddr >> 16) != ((s.tid() >> 5) & 3) * 32) {
        fprintf(stderr, "tc_emu: tcgen05.ld from lanes %u by warp %d (a warp may only read its own lane quadrant)\n",
                taddr >> 16, s.tid() >> 5);
        abort();
    }
    memcpy(v, tmem_of(s.cta()) + (size_t)lane * 512 + col, 32 * sizeof(float));
}
inline void tmem_ld_wait() {}

// ---- tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), operands through shared-memory matrix descriptors --------------------
// descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), SWIZZLE_128B.  Logical (pre-swizzle) byte address of
//   K-major  operand element (mn, k):  start + (mn / 8) * SBO + (mn % 8) * 128 + 2 k
//   MN-major operand element (mn, k):  start + (mn / 64) * LBO + (k / 8) * SBO + (k % 8) * 128 + 2 (mn % 64)
inline float umma_elem(int cta, uint64_t desc, bool mn_major, int mn, int k) {
    const uint32_t start = (uint32_t)(desc & 0x3FFF) << 4, lbo = (uint32_t)((desc >> 16) & 0x3FFF) << 4,
                   sbo = (uint32_t)((desc >> 32) & 0x3FFF) << 4;
    const uint32_t a = mn_major ? start + (mn / 64) * lbo + (k / 8) * sbo + (k % 8) * 128 + 2 * (mn % 64)
                                : start + (mn / 8) * sbo + (mn % 8) * 128 + 2 * k;
    __nv_bfloat16 h;
    memcpy(&h, smem_ptr(((uint32_t)cta << 24) | swz128(a)), 2);
    return __bfloat162float(h);
}
inline void umma_generic(int ctas, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
    const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
    if (M != 128 * ctas || N % ctas) { fprintf(stderr, "tc_emu: MMA shape %dx%d for %d CTA(s)\n", M, N, ctas); abort(); }
    const int col0 = (int)(tmem_d & 0xFFFF), n_per = N / ctas;
    std::vector<float> B((size_t)N * 16);
    for (int n = 0; n < N; ++n)                               // CTA c holds columns [c n_per, (c+1) n_per) of B
        for (int k = 0; k < 16; ++k) B[(size_t)n * 16 + k] = umma_elem(n / n_per, bdesc, b_mn, n % n_per, k);
    for (int c = 0; c < ctas; ++c) {                          // CTA c holds rows [128 c, 128 c + 128) of A and of D
        float* D = tmem_of(c);
        for (int m = 0; m < 128; ++m) {
            float a[16];
            for (int k = 0; k < 16; ++k) a[k] = umma_elem(c, adesc, a_mn, m, k);
            for (int n = 0; n < N; ++n) {
                float acc = 0.f;
                for (int k = 0; k < 16; ++k) acc += a[k] * B[(size_t)n * 16 + k];
                float& d = D[(size_t)m * 512 + col0 + n];
                d = accum ? d + acc : acc;
            }
        }
    }
}
inline void umma_bf16(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum) { umma_generic(1, tmem_d, ad, bd, idesc, accum); }
inline void umma_bf16_pair(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum) { umma_generic(2, tmem_d, ad, bd, idesc, accum); }
// the MMAs above have already retired when the commit is issued
inline void umma_commit(uint32_t bar) { mbar_arrive(bar); }
inline void umma_commit_pair(uint32_t bar) { mbar_arrive(map_to_cta(bar, 0)); mbar_arrive(map_to_cta(bar, 1)); }

}  // namespace
}  // namespace svae
