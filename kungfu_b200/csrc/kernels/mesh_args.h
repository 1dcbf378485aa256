// kungfu_b200 — host-safe definitions shared by the kernel library, the peer mesh and the bindings.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace kf {

constexpr int kMaxRanks = 8;    // one NVSwitch domain
constexpr int kMaxBlocks = 160; // >= 148 SMs; signal pad rows per channel
constexpr int kNumChannels = 8; // independent signal pads (concurrent streams)
// One channel's signal pad: flags[kMaxBlocks][kMaxRanks] (peer-written barrier flags) followed by kDoneWords
// peer-incremented completion counters (split-phase collectives: a body kernel's blocks add 1 to word 0 of every
// rank's pad when their stores have been issued, a one-warp kernel waits for the total). The private per-block
// sequence counters (and the expected completion totals) live in a separate, non-shared allocation of
// kSeqWords words per channel.
constexpr int kDoneWords = 64;
constexpr size_t kFlagWords = (size_t)kMaxBlocks * kMaxRanks;
constexpr size_t kSignalPadBytes = (kFlagWords + kDoneWords) * sizeof(uint32_t);
constexpr int kSeqWords = kMaxBlocks + 16; // [0, kMaxBlocks): per-block barrier sequence; [kMaxBlocks + i]: expected total of done word i

enum DType : int { kF32 = 0, kBF16 = 1, kF16 = 2, kI32 = 3, kI64 = 4, kU8 = 5, kF64 = 6, kI8 = 7, kU32 = 8, kU64 = 9, kI16 = 10, kU16 = 11 };
enum RedOp : int { kSum = 0, kMin = 1, kMax = 2, kProd = 3 };

inline int dtype_size(int dt) {
    switch (dt) {
    case kF32: case kI32: case kU32: return 4;
    case kBF16: case kF16: case kI16: case kU16: return 2;
    case kI64: case kF64: case kU64: return 8;
    default: return 1;
    }
}

struct MeshArgs {
    void* data[kMaxRanks];      // peer-mapped arena bases (data[rank] is the local arena)
    uint32_t* sig[kMaxRanks];   // peer-mapped signal pad bases for the selected channel
    void* mc;                   // multicast (NVLS) alias of the arena, or nullptr
    uint32_t* seq;              // local: per-block sequence counters for the selected channel
    uint32_t* status;           // host-mapped: [0]=error code, [1]=rank/peer, [2]=#timeouts, [3]=want, [4]=seen
    unsigned long long timeout_ns;
    int rank;
    int nranks;
};

enum SignalError : uint32_t { kOk = 0, kErrBarrierTimeout = 1, kErrSeqlockTorn = 2, kErrAbortedByHost = 3 };

} // namespace kf
