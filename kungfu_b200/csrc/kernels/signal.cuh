// kungfu_b200 — cross-GPU flag barriers over peer-mapped signal pads.
//
// Replaces the reference's CPU-side rendezvous (named recv queues, srcs/go/rchannel/handler/
// collective.go:10-65) and its `Barrier` = all-reduce of np bytes (srcs/go/kungfu/session/
// session.go:96-113) for everything that runs inside a kernel.
//
// Protocol. Every rank owns a signal pad per channel: flags[kMaxBlocks][kMaxRanks] (uint32) plus a
// private seq[kMaxBlocks]. Block b on rank r only ever talks to block b on its peers, so no
// grid-wide co-residency is required (a late CTA only delays its peers, it cannot deadlock them).
// A barrier bumps the block's sequence number s, writes s into slot [b][r] of every peer's pad and
// spins until slots [b][*] of the local pad are all >= s (signed-difference compare, so wrap-around
// is harmless). A peer can be at most one barrier ahead, so monotonic numbers never alias and pads
// never need re-zeroing between calls. All ranks must issue the same sequence of (channel, grid)
// launches — guaranteed by the deterministic bucket plan (consensus-checked once on the host).
//
// Failure containment: every spin has a wall-clock budget (globaltimer). On expiry the block writes
// an error code into a host-mapped status word and stops waiting, so a dead peer produces a
// diagnosable error instead of a hung GPU.
#pragma once
#include "common.cuh"

namespace kf {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spin until *p - want >= 0 (signed). Returns false on timeout.
__device__ __forceinline__ bool spin_ge(const uint32_t* p, uint32_t want, const MeshArgs& m, int peer) {
    uint32_t v = ld_acquire_sys(p);
    if ((int32_t)(v - want) >= 0) return true;
    const unsigned long long t0 = global_ns();
    uint32_t iters = 0;
    while (true) {
        v = ld_acquire_sys(p);
        if ((int32_t)(v - want) >= 0) return true;
        if (((++iters) & 0x3ff) == 0) {
            // a wait that already failed once on this mesh poisons every later wait: the job is lost, and each remaining
            // kernel of the step would otherwise sit out its own full time-out (the host raises at its next `check()`). The
            // status word lives in host memory (a PCIe round trip), so it is looked at here — every 1024 polls — and not on
            // the way into the loop, where it cost every barrier ~2 us.
            if (m.status && *((volatile uint32_t*)&m.status[0]) != 0u) return false;
            if (global_ns() - t0 > m.timeout_ns) {
                if (m.status) {
                    m.status[1] = ((uint32_t)m.rank << 16) | (uint32_t)peer;
                    m.status[3] = want;
                    m.status[4] = v;
                    m.status[2] = m.status[2] + 1u;
                    *((volatile uint32_t*)&m.status[0]) = kErrBarrierTimeout;
                    __threadfence_system();
                }
                return false;
            }
            __nanosleep(64);
        }
    }
}

// Per-block barrier with the same block index on every peer. Must be called by all threads of the
// block. `seq` is the block's running sequence number (kept in a register by the caller, loaded
// from m.seq[blockIdx.x] at kernel start and stored back at kernel end).
//   kRelease: my prior writes (incl. writes of other threads of this block, ordered by bar.sync)
//             become visible to peers that observe the flag; and peers' prior writes become visible
//             to me. The start-of-kernel barrier only needs "the peer kernel is running" and uses
//             the relaxed store.
template <bool kRelease> __device__ __forceinline__ void block_barrier(const MeshArgs& m, uint32_t& seq) {
    __syncthreads();
    seq += 1;
    if (threadIdx.x < (unsigned)m.nranks) {
        const int peer = threadIdx.x;
        uint32_t* dst = m.sig[peer] + (size_t)blockIdx.x * kMaxRanks + m.rank;
        if constexpr (kRelease) st_release_sys(dst, seq);
        else st_relaxed_sys(dst, seq);
        const uint32_t* src = m.sig[m.rank] + (size_t)blockIdx.x * kMaxRanks + peer;
        spin_ge(src, seq, m, peer);
    }
    __syncthreads();
}

// ---- split-phase completion (gate kernel -> non-spinning body -> one-warp wait kernel) ----------------------------------
// A body kernel launched between a gate (`barrier_kernel`: every rank has reached this point of its stream) and a wait
// (`wait_done_kernel`) never spins: its CTAs do their loads/stores over NVLink, add 1 to done word `word` of EVERY rank's
// pad and exit, so no SM is parked while a peer is late — the waiting is done by one warp. Called by all threads of the
// block after their last store.
__device__ __forceinline__ void red_release_sys_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void signal_done(const MeshArgs& m, int word) {
    __syncthreads(); // every thread's stores are issued; the release below is cumulative over them
    if (threadIdx.x < (unsigned)m.nranks) {
        __threadfence_system();
        red_release_sys_add(m.sig[threadIdx.x] + kFlagWords + word, 1u);
    }
}
// One thread: wait until `add` more completions have arrived on local done word `word` (running total kept in m.seq).
__device__ __forceinline__ void wait_done(const MeshArgs& m, int word, uint32_t add) {
    const uint32_t want = m.seq[kMaxBlocks + word] + add;
    spin_ge(m.sig[m.rank] + kFlagWords + word, want, m, 0xffff);
    m.seq[kMaxBlocks + word] = want;
}

__device__ __forceinline__ uint32_t load_seq(const MeshArgs& m) { return m.seq[blockIdx.x]; }
__device__ __forceinline__ void store_seq(const MeshArgs& m, uint32_t seq) {
    // all threads of the block passed the final __syncthreads of the last barrier
    if (threadIdx.x == 0) m.seq[blockIdx.x] = seq;
}

} // namespace kf
