// Implementation of peer_mesh.h (replaces the reference NCCLHelper / communicator bootstrap, srcs/cpp/src/nccl/helper.cpp, with
// CUDA VMM export/import + NVLS multicast binding; design notes in the header).
#include "peer_mesh.h"

#include <cuda.h>
#include <errno.h>
#include <poll.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <map>
#include <random>
#include <sstream>
#include <stdexcept>

namespace kf {

#define KF_RT_CHECK(expr)                                                                                               \
    do {                                                                                                                \
        cudaError_t _e = (expr);                                                                                        \
        if (_e != cudaSuccess)                                                                                          \
            throw std::runtime_error(std::string("kungfu_b200 PeerMesh: ") + cudaGetErrorString(_e) + " in " #expr);    \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Driver API through cudaGetDriverEntryPoint: no link-time dependency on libcuda, so the extension
// imports on GPU-less machines (build check, CPU tests).
// ------------------------------------------------------------------------------------------------
namespace {

struct Driver {
    decltype(&cuMemCreate) MemCreate = nullptr;
    decltype(&cuMemRelease) MemRelease = nullptr;
    decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
    decltype(&cuMemAddressFree) MemAddressFree = nullptr;
    decltype(&cuMemMap) MemMap = nullptr;
    decltype(&cuMemUnmap) MemUnmap = nullptr;
    decltype(&cuMemSetAccess) MemSetAccess = nullptr;
    decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
    decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
    decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
    decltype(&cuMulticastCreate) MulticastCreate = nullptr;
    decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
    decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
    decltype(&cuMulticastUnbind) MulticastUnbind = nullptr;
    decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
    decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
    decltype(&cuGetErrorString) GetErrorString = nullptr;
    bool ok = false;
    std::string why;
};

template <typename F> bool load_sym(const char* name, F& fn, std::string& why) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || p == nullptr || q != cudaDriverEntryPointSuccess) {
        why = std::string("driver symbol ") + name + " unavailable";
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

Driver& drv() {
    static Driver d;
    static bool tried = false;
    if (!tried) {
        tried = true;
        bool ok = true;
#define KF_LOAD(sym) ok = ok && load_sym("cu" #sym, d.sym, d.why)
        KF_LOAD(MemCreate);
        KF_LOAD(MemRelease);
        KF_LOAD(MemAddressReserve);
        KF_LOAD(MemAddressFree);
        KF_LOAD(MemMap);
        KF_LOAD(MemUnmap);
        KF_LOAD(MemSetAccess);
        KF_LOAD(MemGetAllocationGranularity);
        KF_LOAD(MemExportToShareableHandle);
        KF_LOAD(MemImportFromShareableHandle);
        KF_LOAD(DeviceGetAttribute);
        KF_LOAD(GetErrorString);
        d.ok = ok;
        // multicast symbols are optional
        std::string w2;
        bool mc = true;
#define KF_LOAD_MC(sym) mc = mc && load_sym("cu" #sym, d.sym, w2)
        KF_LOAD_MC(MulticastCreate);
        KF_LOAD_MC(MulticastAddDevice);
        KF_LOAD_MC(MulticastBindMem);
        KF_LOAD_MC(MulticastUnbind);
        KF_LOAD_MC(MulticastGetGranularity);
        if (!mc) d.MulticastCreate = nullptr;
    }
    return d;
}

std::string cu_err(CUresult r) {
    const char* s = nullptr;
    if (drv().GetErrorString) drv().GetErrorString(r, &s);
    std::ostringstream os;
    os << "CUresult " << (int)r << (s ? std::string(" (") + s + ")" : std::string());
    return os.str();
}

#define KF_CU_CHECK(expr)                                                                                               \
    do {                                                                                                                \
        CUresult _r = (expr);                                                                                           \
        if (_r != CUDA_SUCCESS) throw std::runtime_error(std::string("kungfu_b200 PeerMesh: ") + cu_err(_r) + " in " #expr); \
    } while (0)

size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

// ---- fd passing over abstract Unix datagram sockets ------------------------------------------------
class FdChannel {
public:
    FdChannel(const std::string& prefix, int rank) : prefix_(prefix) {
        sock_ = ::socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
        if (sock_ < 0) throw std::runtime_error("kungfu_b200 FdChannel: socket() failed");
        sockaddr_un a;
        socklen_t len = make_addr(rank, a);
        if (::bind(sock_, reinterpret_cast<sockaddr*>(&a), len) != 0) {
            ::close(sock_);
            throw std::runtime_error(std::string("kungfu_b200 FdChannel: bind failed: ") + strerror(errno));
        }
    }
    ~FdChannel() {
        if (sock_ >= 0) ::close(sock_);
    }
    void send_fd(int to, int fd, int src, int tag) {
        sockaddr_un a;
        socklen_t len = make_addr(to, a);
        int payload[2] = {src, tag};
        iovec iov{payload, sizeof(payload)};
        alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
        memset(ctrl, 0, sizeof(ctrl));
        msghdr msg{};
        msg.msg_name = &a;
        msg.msg_namelen = len;
        msg.msg_iov = &iov;
        msg.msg_iovlen = 1;
        msg.msg_control = ctrl;
        msg.msg_controllen = sizeof(ctrl);
        cmsghdr* c = CMSG_FIRSTHDR(&msg);
        c->cmsg_level = SOL_SOCKET;
        c->cmsg_type = SCM_RIGHTS;
        c->cmsg_len = CMSG_LEN(sizeof(int));
        memcpy(CMSG_DATA(c), &fd, sizeof(int));
        for (int attempt = 0; attempt < 200; ++attempt) {
            if (::sendmsg(sock_, &msg, 0) >= 0) return;
            if (errno == ECONNREFUSED || errno == ENOENT || errno == EAGAIN || errno == ENOBUFS) {
                ::usleep(10000);
                continue;
            }
            break;
        }
        throw std::runtime_error(std::string("kungfu_b200 FdChannel: sendmsg failed: ") + strerror(errno));
    }
    // receive the fd tagged (src, tag); other arrivals are stashed
    int recv_fd(int src, int tag, int timeout_ms) {
        const auto key = std::make_pair(src, tag);
        while (true) {
            auto it = stash_.find(key);
            if (it != stash_.end()) {
                int fd = it->second;
                stash_.erase(it);
                return fd;
            }
            pollfd p{sock_, POLLIN, 0};
            int pr = ::poll(&p, 1, timeout_ms);
            if (pr <= 0) throw std::runtime_error("kungfu_b200 FdChannel: timed out waiting for a peer's descriptor");
            int payload[2] = {-1, -1};
            iovec iov{payload, sizeof(payload)};
            alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
            msghdr msg{};
            msg.msg_iov = &iov;
            msg.msg_iovlen = 1;
            msg.msg_control = ctrl;
            msg.msg_controllen = sizeof(ctrl);
            ssize_t n = ::recvmsg(sock_, &msg, MSG_CMSG_CLOEXEC);
            if (n < (ssize_t)sizeof(payload)) continue;
            cmsghdr* c = CMSG_FIRSTHDR(&msg);
            if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) continue;
            int fd = -1;
            memcpy(&fd, CMSG_DATA(c), sizeof(int));
            stash_[std::make_pair(payload[0], payload[1])] = fd;
        }
    }

private:
    socklen_t make_addr(int rank, sockaddr_un& a) const {
        memset(&a, 0, sizeof(a));
        a.sun_family = AF_UNIX;
        std::string name = prefix_ + "-" + std::to_string(rank);
        if (name.size() + 1 > sizeof(a.sun_path)) name.resize(sizeof(a.sun_path) - 1);
        a.sun_path[0] = '\0'; // abstract namespace: no filesystem residue
        memcpy(a.sun_path + 1, name.data(), name.size());
        return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
    }
    int sock_ = -1;
    std::string prefix_;
    std::map<std::pair<int, int>, int> stash_;
};

// all ranks report ok/fail; true iff everyone succeeded
bool agree(const OobAllGather& oob, bool ok) {
    auto all = oob(ok ? "1" : "0");
    for (auto& s : all)
        if (s != "1") return false;
    return true;
}

} // namespace

// ------------------------------------------------------------------------------------------------
PeerMesh::PeerMesh(int device, size_t arena_bytes, const std::string& mode) : device_(device), mode_(mode) {
    KF_RT_CHECK(cudaSetDevice(device_));
    KF_RT_CHECK(cudaFree(nullptr));
    arena_bytes_ = round_up(arena_bytes, 4096);
    signal_off_ = arena_bytes_;
    const size_t want = arena_bytes_ + (size_t)kNumChannels * kSignalPadBytes;
    alloc_bytes_ = want;
    if (mode_ == "auto") mode_ = drv().ok ? "vmm" : "ipc";
    if (mode_ == "vmm" && !drv().ok) throw std::runtime_error("kungfu_b200 PeerMesh: VMM driver API unavailable: " + drv().why);
    alloc_local();
    KF_RT_CHECK(cudaMalloc(&seq_, sizeof(uint32_t) * kNumChannels * kSeqWords));
    KF_RT_CHECK(cudaMemset(seq_, 0, sizeof(uint32_t) * kNumChannels * kSeqWords));
    KF_RT_CHECK(cudaHostAlloc(&status_host_, 64 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(status_host_, 0, 64 * sizeof(uint32_t));
    KF_RT_CHECK(cudaHostGetDevicePointer(&status_dev_, status_host_, 0));
    peer_va_.assign(1, local_va_);
    peer_handle_.assign(1, 0);
    peer_is_ipc_.assign(1, false);
    KF_RT_CHECK(cudaDeviceSynchronize());
}

PeerMesh::~PeerMesh() {
    try {
        cudaSetDevice(device_);
        cudaDeviceSynchronize();
        disconnect();
        free_local();
        if (seq_) cudaFree(seq_);
        if (status_host_) cudaFreeHost(status_host_);
    } catch (...) {
    }
}

void PeerMesh::alloc_local() {
    if (mode_ == "vmm") {
        Driver& d = drv();
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = device_;
        prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        size_t gran = 0;
        KF_CU_CHECK(d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
        if (gran == 0) gran = 2u << 20;
        size_t g = gran;
        if (d.MulticastCreate) {
            // size the allocation so that it can be bound to a multicast object later
            CUmulticastObjectProp mp{};
            mp.numDevices = 2;
            mp.size = round_up(alloc_bytes_, gran);
            mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
            size_t mg = 0;
            if (d.MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
        }
        alloc_bytes_ = round_up(alloc_bytes_, g);
        CUmemGenericAllocationHandle h;
        KF_CU_CHECK(d.MemCreate(&h, alloc_bytes_, &prop, 0));
        CUdeviceptr va = 0;
        CUresult r = d.MemAddressReserve(&va, alloc_bytes_, g, 0, 0);
        if (r != CUDA_SUCCESS) {
            d.MemRelease(h);
            throw std::runtime_error("kungfu_b200 PeerMesh: cuMemAddressReserve failed: " + cu_err(r));
        }
        KF_CU_CHECK(d.MemMap(va, alloc_bytes_, 0, h, 0));
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device_;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        KF_CU_CHECK(d.MemSetAccess(va, alloc_bytes_, &acc, 1));
        local_va_ = reinterpret_cast<void*>(va);
        local_handle_ = (unsigned long long)h;
        local_is_vmm_ = true;
    } else {
        alloc_bytes_ = round_up(alloc_bytes_, 2u << 20);
        KF_RT_CHECK(cudaMalloc(&local_va_, alloc_bytes_));
        local_is_vmm_ = false;
    }
    KF_RT_CHECK(cudaMemset(local_va_, 0, alloc_bytes_));
}

void PeerMesh::free_local() {
    if (!local_va_) return;
    if (local_is_vmm_) {
        Driver& d = drv();
        d.MemUnmap((CUdeviceptr)local_va_, alloc_bytes_);
        d.MemAddressFree((CUdeviceptr)local_va_, alloc_bytes_);
        d.MemRelease((CUmemGenericAllocationHandle)local_handle_);
    } else {
        cudaFree(local_va_);
    }
    local_va_ = nullptr;
}

void PeerMesh::disconnect() {
    if (!connected_) return;
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    Driver& d = drv();
    if (mc_va_) {
        d.MemUnmap((CUdeviceptr)mc_va_, alloc_bytes_);
        d.MemAddressFree((CUdeviceptr)mc_va_, alloc_bytes_);
        mc_va_ = nullptr;
    }
    if (mc_handle_) {
        if (mc_bound_ && d.MulticastUnbind) d.MulticastUnbind((CUmemGenericAllocationHandle)mc_handle_, device_, 0, alloc_bytes_);
        d.MemRelease((CUmemGenericAllocationHandle)mc_handle_);
        mc_handle_ = 0;
        mc_bound_ = false;
    }
    for (int r = 0; r < (int)peer_va_.size(); ++r) {
        if (r == rank_ || peer_va_[r] == nullptr || peer_va_[r] == local_va_) continue;
        if (peer_is_ipc_[r]) {
            cudaIpcCloseMemHandle(peer_va_[r]);
        } else {
            d.MemUnmap((CUdeviceptr)peer_va_[r], alloc_bytes_);
            d.MemAddressFree((CUdeviceptr)peer_va_[r], alloc_bytes_);
            d.MemRelease((CUmemGenericAllocationHandle)peer_handle_[r]);
        }
    }
    peer_va_.assign(1, local_va_);
    peer_handle_.assign(1, 0);
    peer_is_ipc_.assign(1, false);
    rank_ = 0;
    nranks_ = 1;
    connected_ = false;
}

void PeerMesh::reset_signals(cudaStream_t st) {
    KF_RT_CHECK(cudaMemsetAsync(reinterpret_cast<char*>(local_va_) + signal_off_, 0, (size_t)kNumChannels * kSignalPadBytes, st));
    KF_RT_CHECK(cudaMemsetAsync(seq_, 0, sizeof(uint32_t) * kNumChannels * kSeqWords, st));
    KF_RT_CHECK(cudaStreamSynchronize(st));
}

void PeerMesh::connect(int rank, int nranks, const OobAllGather& oob, bool want_multicast) {
    if (nranks < 1 || nranks > kMaxRanks) throw std::runtime_error("kungfu_b200 PeerMesh: world size must be 1.." + std::to_string(kMaxRanks));
    if (rank < 0 || rank >= nranks) throw std::runtime_error("kungfu_b200 PeerMesh: bad rank");
    KF_RT_CHECK(cudaSetDevice(device_));
    disconnect();
    epoch_ += 1;
    rank_ = rank;
    nranks_ = nranks;
    mc_note_.clear();
    // stale flags from a previous membership must be gone before any peer can signal again
    reset_signals(nullptr);
    peer_va_.assign(nranks, nullptr);
    peer_handle_.assign(nranks, 0);
    peer_is_ipc_.assign(nranks, false);
    peer_va_[rank] = local_va_;
    if (nranks > 1) {
        // every rank must agree on allocation size and mode
        std::ostringstream me;
        me << alloc_bytes_ << ":" << mode_;
        auto all = oob(me.str());
        for (auto& s : all)
            if (s != me.str())
                throw std::runtime_error("kungfu_b200 PeerMesh: ranks disagree on arena size/mode: " + s + " vs " + me.str());
        if (mode_ == "vmm") connect_vmm(oob, want_multicast);
        else if (mode_ == "ipc") connect_ipc(oob);
        else throw std::runtime_error("kungfu_b200 PeerMesh: mode '" + mode_ + "' cannot span ranks");
        oob("ready");
    } else if (want_multicast) {
        mc_note_ = "single rank: multicast not needed";
    }
    connected_ = true;
}

void PeerMesh::connect_vmm(const OobAllGather& oob, bool want_multicast) {
    Driver& d = drv();
    // job token chosen by rank 0
    std::random_device rd;
    std::ostringstream tok;
    tok << "kfb200-" << ::getpid() << "-" << rd() << "-e" << epoch_;
    const std::string prefix = oob(tok.str()).at(0);
    FdChannel chan(prefix, rank_);
    oob("bound");

    int fd = -1;
    KF_CU_CHECK(d.MemExportToShareableHandle(&fd, (CUmemGenericAllocationHandle)local_handle_,
                                             CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    for (int r = 0; r < nranks_; ++r)
        if (r != rank_) chan.send_fd(r, fd, rank_, 0);
    ::close(fd);

    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    for (int r = 0; r < nranks_; ++r) {
        if (r == rank_) continue;
        int pfd = chan.recv_fd(r, 0, 60000);
        CUmemGenericAllocationHandle h;
        CUresult res = d.MemImportFromShareableHandle(&h, (void*)(uintptr_t)pfd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        ::close(pfd);
        if (res != CUDA_SUCCESS) throw std::runtime_error("kungfu_b200 PeerMesh: import of peer allocation failed: " + cu_err(res));
        CUdeviceptr va = 0;
        KF_CU_CHECK(d.MemAddressReserve(&va, alloc_bytes_, 0, 0, 0));
        KF_CU_CHECK(d.MemMap(va, alloc_bytes_, 0, h, 0));
        KF_CU_CHECK(d.MemSetAccess(va, alloc_bytes_, &acc, 1));
        peer_va_[r] = reinterpret_cast<void*>(va);
        peer_handle_[r] = (unsigned long long)h;
    }
    oob("mapped");

    // ---- NVLS multicast (optional) ----
    if (!want_multicast) {
        mc_note_ = "multicast not requested";
        return;
    }
    int supported = 0;
    if (d.MulticastCreate) d.DeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device_);
    if (!agree(oob, supported != 0)) {
        mc_note_ = "multicast unsupported on at least one rank";
        return;
    }
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)nranks_;
    mp.size = alloc_bytes_;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mp.flags = 0;
    bool ok = true;
    std::string why;
    CUmemGenericAllocationHandle mc = 0;
    if (rank_ == 0) {
        CUresult r = d.MulticastCreate(&mc, &mp);
        int mfd = -1;
        if (r == CUDA_SUCCESS) r = d.MemExportToShareableHandle(&mfd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
        if (r != CUDA_SUCCESS) {
            ok = false;
            why = "create/export: " + cu_err(r);
            if (mc) d.MemRelease(mc), mc = 0;
        }
        // always send something so that peers do not block: -1 cannot be sent, so send our own arena fd on failure
        if (!agree(oob, ok)) {
            mc_note_ = "multicast create failed: " + why;
            if (mfd >= 0) ::close(mfd);
            if (mc) d.MemRelease(mc);
            return;
        }
        for (int r2 = 1; r2 < nranks_; ++r2) chan.send_fd(r2, mfd, 0, 1);
        ::close(mfd);
    } else {
        if (!agree(oob, true)) {
            mc_note_ = "multicast create failed on rank 0";
            return;
        }
        int mfd = chan.recv_fd(0, 1, 60000);
        CUresult r = d.MemImportFromShareableHandle(&mc, (void*)(uintptr_t)mfd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        ::close(mfd);
        if (r != CUDA_SUCCESS) {
            ok = false;
            why = "import: " + cu_err(r);
            mc = 0;
        }
    }
    if (ok) {
        CUresult r = d.MulticastAddDevice(mc, device_);
        if (r != CUDA_SUCCESS) {
            ok = false;
            why = "add device: " + cu_err(r);
        }
    }
    if (!agree(oob, ok)) {
        mc_note_ = "multicast setup failed: " + why;
        if (mc) d.MemRelease(mc);
        return;
    }
    // all devices have been added; now bind the local allocation and map the multicast alias
    CUdeviceptr mva = 0;
    bool bound = false;
    {
        CUresult r = d.MulticastBindMem(mc, 0, (CUmemGenericAllocationHandle)local_handle_, 0, alloc_bytes_, 0);
        if (r == CUDA_SUCCESS) {
            bound = true;
            r = d.MemAddressReserve(&mva, alloc_bytes_, 0, 0, 0);
            if (r == CUDA_SUCCESS) {
                r = d.MemMap(mva, alloc_bytes_, 0, mc, 0);
                if (r == CUDA_SUCCESS) r = d.MemSetAccess(mva, alloc_bytes_, &acc, 1);
                if (r != CUDA_SUCCESS) {
                    d.MemAddressFree(mva, alloc_bytes_);
                    mva = 0;
                }
            }
        }
        if (r != CUDA_SUCCESS) {
            ok = false;
            why = "bind/map: " + cu_err(r);
        }
    }
    if (!agree(oob, ok)) {
        mc_note_ = "multicast bind failed: " + why;
        if (mva) {
            d.MemUnmap(mva, alloc_bytes_);
            d.MemAddressFree(mva, alloc_bytes_);
        }
        if (bound && d.MulticastUnbind) d.MulticastUnbind(mc, device_, 0, alloc_bytes_);
        d.MemRelease(mc);
        return;
    }
    mc_handle_ = (unsigned long long)mc;
    mc_bound_ = true;
    mc_va_ = reinterpret_cast<void*>(mva);
    mc_note_ = "multicast bound";
}

void PeerMesh::connect_ipc(const OobAllGather& oob) {
    cudaIpcMemHandle_t h;
    KF_RT_CHECK(cudaIpcGetMemHandle(&h, local_va_));
    std::string mine(reinterpret_cast<const char*>(&h), sizeof(h));
    auto all = oob(mine);
    for (int r = 0; r < nranks_; ++r) {
        if (r == rank_) continue;
        if (all[r].size() != sizeof(h)) throw std::runtime_error("kungfu_b200 PeerMesh: bad IPC handle from peer");
        cudaIpcMemHandle_t ph;
        memcpy(&ph, all[r].data(), sizeof(ph));
        void* p = nullptr;
        KF_RT_CHECK(cudaIpcOpenMemHandle(&p, ph, cudaIpcMemLazyEnablePeerAccess));
        peer_va_[r] = p;
        peer_is_ipc_[r] = true;
    }
    mc_note_ = "ipc mode: no multicast";
}

MeshArgs PeerMesh::args(int channel) const {
    if (channel < 0 || channel >= kNumChannels) throw std::runtime_error("kungfu_b200 PeerMesh: bad channel");
    MeshArgs m{};
    for (int r = 0; r < kMaxRanks; ++r) {
        void* base = r < nranks_ ? peer_va_[r] : nullptr;
        m.data[r] = base;
        m.sig[r] = base ? reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(base) + signal_off_ + (size_t)channel * kSignalPadBytes)
                        : nullptr;
    }
    m.mc = mc_va_;
    m.seq = seq_ + (size_t)channel * kSeqWords;
    m.status = status_dev_;
    m.timeout_ns = timeout_ns_;
    m.rank = rank_;
    m.nranks = nranks_;
    return m;
}

std::vector<uint32_t> PeerMesh::status() const {
    std::vector<uint32_t> v(8);
    for (int i = 0; i < 8; ++i) v[i] = ((volatile uint32_t*)status_host_)[i];
    return v;
}
void PeerMesh::clear_status() { memset(status_host_, 0, 64 * sizeof(uint32_t)); }
void PeerMesh::abort() {
    if (!status_host_) return;
    volatile uint32_t* st = (volatile uint32_t*)status_host_;
    if (st[0] == 0) {
        st[1] = ((uint32_t)rank_ << 16) | 0xffffu;
        __sync_synchronize();
        st[0] = kErrAbortedByHost;
    }
}

MeshTopology PeerMesh::topology() const {
    MeshTopology t;
    t.device = device_;
    cudaGetDeviceCount(&t.device_count);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device_) == cudaSuccess) {
        t.name = p.name;
        t.sm_count = p.multiProcessorCount;
        t.total_mem = p.totalGlobalMem;
    }
    for (int d2 = 0; d2 < t.device_count; ++d2) {
        int can = d2 == device_ ? 1 : 0, perf = 0, atom = 0;
        if (d2 != device_) {
            cudaDeviceCanAccessPeer(&can, device_, d2);
            cudaDeviceGetP2PAttribute(&perf, cudaDevP2PAttrPerformanceRank, device_, d2);
            cudaDeviceGetP2PAttribute(&atom, cudaDevP2PAttrNativeAtomicSupported, device_, d2);
        }
        t.can_access.push_back(can);
        t.perf_rank.push_back(perf);
        t.native_atomic.push_back(atom);
    }
    if (drv().ok && drv().MulticastCreate) {
        int s = 0;
        if (drv().DeviceGetAttribute(&s, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device_) == CUDA_SUCCESS) t.multicast_supported = s;
    }
    cudaGetLastError();
    return t;
}

} // namespace kf
