// kungfu_b200 runtime — host blob stores backing P2P requests (PairAveraging's "model store" on CPU).
//
// Reference: srcs/go/store/{blob.go:10-26, store.go:14-59, versionedstore.go:8-94}: name -> blob with a RW lock
// (`GetOrCreate` fails with a conflict if the size differs), and a versioned store keeping a sliding window of the
// last 3 versions (srcs/go/rchannel/handler/p2p.go:11).
#pragma once
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

namespace kfrt {

struct Blob {
    std::shared_mutex mu;
    std::vector<uint8_t> data;
};

class Store {
public:
    // create (or overwrite the content of) blob `name`; throws kfrt::Error on a size conflict
    void save(const std::string& name, const void* data, size_t len);
    std::shared_ptr<Blob> get(const std::string& name);             // nullptr if absent
    std::shared_ptr<Blob> get_or_create(const std::string& name, size_t len);
    bool erase(const std::string& name);
    size_t size();
    std::vector<std::string> names();

private:
    std::mutex mu_;
    std::map<std::string, std::shared_ptr<Blob>> blobs_;
};

class VersionedStore {
public:
    explicit VersionedStore(int window) : window_(window) {}
    void save(const std::string& version, const std::string& name, const void* data, size_t len);
    std::shared_ptr<Blob> get(const std::string& version, const std::string& name); // nullptr if absent
    std::vector<std::string> versions();

private:
    int window_;
    std::mutex mu_;
    std::deque<std::pair<std::string, std::shared_ptr<Store>>> vs_;
};

} // namespace kfrt
