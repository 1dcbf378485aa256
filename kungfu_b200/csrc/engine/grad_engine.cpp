// Implementation of grad_engine.h (design notes and reference mapping there).
#include "grad_engine.h"

#include <stdexcept>

namespace kf {

#define KF_ENG_CHECK(expr)                                                                                              \
    do {                                                                                                                \
        cudaError_t _e = (expr);                                                                                        \
        if (_e != cudaSuccess) throw std::runtime_error(std::string("kungfu_b200 GradEngine: ") + cudaGetErrorString(_e) + " in " #expr); \
    } while (0)

GradEngine::GradEngine(PeerMesh* mesh, int mode, bool average, bool gated, int max_blocks, int threads, int algo)
    : mesh_(mesh), mode_(mode), average_(average), gated_(gated), max_blocks_(max_blocks), threads_(threads), algo_(algo) {
    if (!mesh) throw std::runtime_error("kungfu_b200 GradEngine: no mesh");
    if (mode != kModeSsgd && mode != kModeGrads && mode != kModeCopy) throw std::runtime_error("kungfu_b200 GradEngine: bad mode");
}

GradEngine::~GradEngine() {
    for (auto e : events_) cudaEventDestroy(e);
    if (own_stream_ && comm_) cudaStreamDestroy(comm_);
}

int GradEngine::add_set(size_t g_off, size_t p_off, int dtype, uintptr_t hyper, int opt_kind) {
    EngineSet s;
    s.g_off = g_off;
    s.p_off = p_off;
    s.dtype = dtype;
    s.hyper = reinterpret_cast<float*>(hyper);
    s.opt_kind = opt_kind;
    sets_.push_back(std::move(s));
    return (int)sets_.size() - 1;
}

int GradEngine::add_bucket(int set, size_t start_bytes, size_t nbytes, size_t numel) {
    EngineBucket b;
    b.start_bytes = start_bytes;
    b.nbytes = nbytes;
    b.numel = numel;
    sets_.at(set).buckets.push_back(std::move(b));
    return (int)sets_[set].buckets.size() - 1;
}

int GradEngine::add_entry(int set, int bucket, size_t offset_bytes, size_t nbytes) {
    EngineSet& s = sets_.at(set);
    EngineEntry e;
    e.offset_bytes = offset_bytes;
    e.nbytes = nbytes;
    e.bucket = bucket;
    s.entries.push_back(e);
    s.buckets.at(bucket).n_entries += 1;
    return (int)s.entries.size() - 1;
}

void GradEngine::set_bucket_state(int set, int bucket, uintptr_t master, uintptr_t s1, uintptr_t s2) {
    EngineBucket& b = sets_.at(set).buckets.at(bucket);
    b.master = reinterpret_cast<float*>(master);
    b.s1 = reinterpret_cast<float*>(s1);
    b.s2 = reinterpret_cast<float*>(s2);
}

cudaEvent_t GradEngine::next_event() {
    // a small ring: an event is re-recorded only after everything that waited on its previous recording has been enqueued
    if (events_.size() < 64) {
        cudaEvent_t e;
        KF_ENG_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        events_.push_back(e);
        return e;
    }
    return events_[ev_next_++ % events_.size()];
}

int GradEngine::pick_blocks(size_t nbytes) const {
    const int n = mesh_->nranks();
    const size_t nvec = (nbytes + 15) / 16;
    const size_t work = (nvec + n - 1) / n;
    const size_t per_block = (size_t)threads_ * 2;
    size_t b = (work + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > (size_t)max_blocks_) b = max_blocks_;
    return (int)b;
}

int GradEngine::pick_algo(int dtype) const {
    if (algo_ == kAlgoTwoShot) return kAlgoTwoShot;
    if (mesh_->has_multicast() && nvls_supported(dtype, kSum)) return kAlgoNvls;
    return kAlgoTwoShot;
}

void GradEngine::open_step(cudaStream_t stream) {
    if (step_open_) return;
    step_open_ = true;
    if (!comm_) {
        int lo = 0, hi = 0;
        KF_ENG_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        KF_ENG_CHECK(cudaStreamCreateWithPriority(&comm_, cudaStreamNonBlocking, hi));
        own_stream_ = true;
    }
    cudaEvent_t e = next_event();
    KF_ENG_CHECK(cudaEventRecord(e, stream));
    KF_ENG_CHECK(cudaStreamWaitEvent(comm_, e, 0));
    for (auto& s : sets_) {
        if (s.hyper && mode_ != kModeGrads) {
            launch_hyper_tick(s.hyper, comm_);
            launches_ += 1;
        }
    }
}

bool GradEngine::grad_ready(int set, int entry, const void* grad_ptr, cudaStream_t stream) {
    if (deferred_) return false;
    EngineSet& s = sets_.at(set);
    EngineEntry& e = s.entries.at(entry);
    if (e.got)
        throw std::runtime_error(
            "kungfu_b200: a second backward() reached a parameter whose bucket was already reduced (and, on the fused path, applied) "
            "in this step. Wrap all but the last backward of an accumulation loop in `with optimizer.no_sync():`, or build the "
            "optimizer with overlap=False.");
    e.got = true;
    EngineBucket& b = s.buckets[e.bucket];
    char* slot = reinterpret_cast<char*>(mesh_->local_ptr()) + s.g_off + e.offset_bytes;
    bool stolen = false;
    if (grad_ptr != nullptr && grad_ptr != slot) {
        b.src.push_back(grad_ptr);
        b.dst.push_back(slot);
        b.bytes.push_back(e.nbytes);
        stolen = true;
    }
    b.ready += 1;
    // strictly in bucket order: every rank issues the same kernel sequence
    while (s.launched < (int)s.buckets.size() && s.buckets[s.launched].ready >= s.buckets[s.launched].n_entries) {
        launch_bucket(s, s.launched, stream, true);
        s.launched += 1;
    }
    return stolen;
}

void GradEngine::launch_bucket(EngineSet& s, int bi, cudaStream_t stream, bool after_stream) {
    open_step(stream);
    EngineBucket& b = s.buckets[bi];
    // monitor: when EVERY gradient of the bucket arrives through the hand-over copy, that copy computes sum |g_local|^2 on the way
    // and the collective body skips its extra pass over the local message (the 3 % the noise-scale monitor used to cost)
    const bool pre_in_copy = sumsq_ != nullptr && mesh_->nranks() > 1 && (int)b.src.size() == b.n_entries;
    if (!b.src.empty()) {
        launch_multi_copy(b.src.data(), b.dst.data(), b.bytes.data(), (int)b.src.size(), stream, pre_in_copy ? sumsq_ : nullptr, s.dtype);
        launches_ += (b.src.size() + kMaxCopyBatch - 1) / kMaxCopyBatch;
        b.src.clear();
        b.dst.clear();
        b.bytes.clear();
        after_stream = true;
    }
    if (mode_ == kModeCopy) return; // the optimizer's own kernel reads the arena on `stream` after the last hand-over
    if (after_stream) {
        cudaEvent_t e = next_event();
        KF_ENG_CHECK(cudaEventRecord(e, stream));
        KF_ENG_CHECK(cudaStreamWaitEvent(comm_, e, 0));
    }
    const int n = mesh_->nranks();
    const MeshArgs m = mesh_->args(0);
    char* base = reinterpret_cast<char*>(mesh_->local_ptr());
    const int flags = sumsq_ ? (kArSumSq | (pre_in_copy ? kArSumSqPost : 0)) : 0;
    if (mode_ == kModeSsgd) {
        if (n == 1) {
            launch_flat_opt(base + s.p_off + b.start_bytes, s.dtype, base + s.g_off + b.start_bytes, s.dtype, b.master, b.s1, b.s2, s.hyper, s.opt_kind,
                            b.numel, comm_);
            launches_ += 1;
            if (sumsq_) {
                launch_sumsq(base + s.g_off + b.start_bytes, s.dtype, b.numel, sumsq_, comm_);
                launch_sumsq(base + s.g_off + b.start_bytes, s.dtype, b.numel, sumsq_ + 1, comm_);
                launches_ += 2;
            }
            return;
        }
        const int blocks = pick_blocks(b.nbytes);
        const int algo = pick_algo(s.dtype);
        if (gated_) launch_barrier(m, comm_);
        launch_ssgd_fused(algo, m, s.g_off + b.start_bytes, s.p_off + b.start_bytes, s.dtype, b.nbytes, b.master, b.s1, b.s2, s.hyper, s.opt_kind,
                          sumsq_, flags | (gated_ ? kArGated : 0), blocks, threads_, comm_);
        if (gated_) launch_wait_done(m, 0, (uint32_t)(n * blocks), comm_);
        launches_ += gated_ ? 3 : 1;
    } else {
        if (n == 1) {
            if (sumsq_) {
                launch_sumsq(base + s.g_off + b.start_bytes, s.dtype, b.numel, sumsq_, comm_);
                launch_sumsq(base + s.g_off + b.start_bytes, s.dtype, b.numel, sumsq_ + 1, comm_);
                launches_ += 2;
            }
            return;
        }
        const int blocks = pick_blocks(b.nbytes);
        const int algo = pick_algo(s.dtype);
        const float scale = average_ ? 1.f / (float)n : 1.f;
        if (gated_) launch_barrier(m, comm_);
        launch_allreduce(algo, m, s.g_off + b.start_bytes, nullptr, b.nbytes, s.dtype, kSum, scale, sumsq_,
                         flags | (average_ ? kArScale : 0) | (gated_ ? kArGated : 0), blocks, threads_, comm_);
        if (gated_) launch_wait_done(m, 0, (uint32_t)(n * blocks), comm_);
        launches_ += gated_ ? 3 : 1;
    }
}

void GradEngine::flush(cudaStream_t stream, bool zero_missing) {
    for (auto& s : sets_) {
        // hand-over mode (gradients arrive through grad_ready): a parameter that received no gradient this step contributes zeros
        char* base = reinterpret_cast<char*>(mesh_->local_ptr()) + s.g_off;
        for (auto& e : s.entries) {
            if (zero_missing && !e.got && s.launched <= e.bucket) KF_ENG_CHECK(cudaMemsetAsync(base + e.offset_bytes, 0, e.nbytes, stream));
        }
        bool first = true;
        while (s.launched < (int)s.buckets.size()) {
            launch_bucket(s, s.launched, stream, first);
            first = false;
            s.launched += 1;
        }
    }
}

uint32_t GradEngine::finish(cudaStream_t stream) {
    if (comm_) {
        cudaEvent_t e = next_event();
        KF_ENG_CHECK(cudaEventRecord(e, comm_));
        KF_ENG_CHECK(cudaStreamWaitEvent(stream, e, 0));
    }
    step_open_ = false;
    for (auto& s : sets_) {
        s.launched = 0;
        for (auto& b : s.buckets) b.ready = 0;
        for (auto& e : s.entries) e.got = false;
    }
    return mesh_->nranks() > 1 ? mesh_->status()[0] : 0u;
}

} // namespace kf
