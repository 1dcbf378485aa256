// Implementation of store.hpp (the header carries the design notes and the reference file:line mapping).
#include "store.hpp"

#include <cstring>

#include "base.hpp"

namespace kfrt {

std::shared_ptr<Blob> Store::get(const std::string& name) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = blobs_.find(name);
    return it == blobs_.end() ? nullptr : it->second;
}

std::shared_ptr<Blob> Store::get_or_create(const std::string& name, size_t len) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = blobs_.find(name);
    if (it != blobs_.end()) {
        if (it->second->data.size() != len)
            fail("store: size conflict for blob '" + name + "': " + std::to_string(it->second->data.size()) + " != " + std::to_string(len));
        return it->second;
    }
    auto b = std::make_shared<Blob>();
    b->data.resize(len);
    blobs_[name] = b;
    return b;
}

void Store::save(const std::string& name, const void* data, size_t len) {
    auto b = get_or_create(name, len);
    std::unique_lock<std::shared_mutex> w(b->mu);
    if (len) memcpy(b->data.data(), data, len);
}

bool Store::erase(const std::string& name) {
    std::lock_guard<std::mutex> g(mu_);
    return blobs_.erase(name) > 0;
}
size_t Store::size() {
    std::lock_guard<std::mutex> g(mu_);
    return blobs_.size();
}
std::vector<std::string> Store::names() {
    std::lock_guard<std::mutex> g(mu_);
    std::vector<std::string> v;
    for (auto& kv : blobs_) v.push_back(kv.first);
    return v;
}

void VersionedStore::save(const std::string& version, const std::string& name, const void* data, size_t len) {
    std::shared_ptr<Store> s;
    {
        std::lock_guard<std::mutex> g(mu_);
        for (auto& kv : vs_)
            if (kv.first == version) s = kv.second;
        if (!s) {
            s = std::make_shared<Store>();
            vs_.emplace_back(version, s);
            while ((int)vs_.size() > window_) vs_.pop_front(); // sliding GC window
        }
    }
    s->save(name, data, len);
}

std::shared_ptr<Blob> VersionedStore::get(const std::string& version, const std::string& name) {
    std::shared_ptr<Store> s;
    {
        std::lock_guard<std::mutex> g(mu_);
        for (auto& kv : vs_)
            if (kv.first == version) s = kv.second;
    }
    return s ? s->get(name) : nullptr;
}

std::vector<std::string> VersionedStore::versions() {
    std::lock_guard<std::mutex> g(mu_);
    std::vector<std::string> v;
    for (auto& kv : vs_) v.push_back(kv.first);
    return v;
}

} // namespace kfrt
