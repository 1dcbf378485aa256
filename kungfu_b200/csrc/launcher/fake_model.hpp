// Gradient-tensor size lists of the "fake models" used by the communication-only tools (no compute, just the tensors a
// data-parallel step would exchange).
//
// Reference: tests/go/fakemodel/{fakemodel,resnet50-imagenet,vgg16-imagenet,slp-mnist,bert}.go and
// tests/cpp/integration/{resnet50_info,vgg_info,bert}.hpp. The lists here are derived from the layer shapes of the torchvision /
// BERT-base architectures instead of being tabulated.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace kftools {

inline std::vector<std::string> fake_model_names() { return {"resnet50", "resnet50-keras", "vgg16", "bert", "bert-base", "bert-large", "slp"}; }

inline std::vector<size_t> fake_model_sizes(const std::string& name);

inline std::vector<size_t> fake_model_sizes(const std::string& name) {
    if (name == "vgg16")
        return {1728, 64, 36864, 64, 73728, 128, 147456, 128, 294912, 256, 589824, 256, 589824, 256, 1179648, 512, 2359296, 512, 2359296, 512,
                2359296, 512, 2359296, 512, 2359296, 512, 102760448, 4096, 16777216, 4096, 4096000, 1000};
    if (name == "slp") return {7840, 10};
    if (name == "bert" || name == "bert-base" || name == "bert-large") {
        // embeddings (word 30522, position 512, type 2) + LayerNorm, L encoder layers, pooler. "bert" is the reference's list: BERT-base
        // with the two Adam slots of every variable plus the optimizer's two scalars (599 tensors, 327.3 M elements); "bert-base" /
        // "bert-large" are the bare gradients of the two model sizes (199 tensors / 109.1 M, 391 tensors / 334.1 M).
        const bool large = name == "bert-large";
        const size_t h = large ? 1024 : 768, ff = 4 * h;
        const int layers = large ? 24 : 12;
        std::vector<size_t> v = {30522 * h, 512 * h, 2 * h, h, h};
        for (int l = 0; l < layers; ++l)
            v.insert(v.end(), {h * h, h, h * h, h, h * h, h, h * h, h, h, h, h * ff, ff, ff * h, h, h, h});
        v.insert(v.end(), {h * h, h});
        if (name == "bert") {
            std::vector<size_t> with_slots;
            for (size_t s : v) with_slots.insert(with_slots.end(), {s, s, s});
            with_slots.insert(with_slots.end(), {1, 1});
            return with_slots;
        }
        return v;
    }
    if (name == "resnet50-keras") {
        // the reference's ResNet-50 list comes from tf.keras, whose convolutions carry a bias: 214 tensors / 25.58 M elements
        std::vector<size_t> v = fake_model_sizes("resnet50");
        std::vector<size_t> biases = {64};
        const size_t mids[4] = {64, 128, 256, 512};
        const int reps[4] = {3, 4, 6, 3};
        for (int s = 0; s < 4; ++s)
            for (int b = 0; b < reps[s]; ++b) {
                biases.insert(biases.end(), {mids[s], mids[s], mids[s] * 4});
                if (b == 0) biases.push_back(mids[s] * 4);
            }
        v.insert(v.end(), biases.begin(), biases.end());
        return v;
    }
    if (name != "resnet50") return {};
    // 161 tensors, 25.5 M elements: stem, 16 bottleneck blocks (1x1, 3x3, 1x1 + their BN pairs, down-sample on the first of a stage), fc
    std::vector<size_t> v = {9408, 64, 64};
    auto block = [&](size_t in, size_t mid, size_t out, bool down) {
        v.insert(v.end(), {in * mid, mid, mid, mid * mid * 9, mid, mid, mid * out, out, out});
        if (down) v.insert(v.end(), {in * out, out, out});
    };
    size_t in = 64;
    const size_t mids[4] = {64, 128, 256, 512};
    const int reps[4] = {3, 4, 6, 3};
    for (int s = 0; s < 4; ++s)
        for (int b = 0; b < reps[s]; ++b) {
            block(in, mids[s], mids[s] * 4, b == 0);
            in = mids[s] * 4;
        }
    v.insert(v.end(), {2048 * 1000, 1000});
    return v;
}

inline size_t total_elements(const std::vector<size_t>& sizes) {
    size_t t = 0;
    for (size_t s : sizes) t += s;
    return t;
}

// "-flag value" / "-flag=value" / boolean "-flag" command lines of the Go tools
struct Flags {
    int argc;
    char** argv;
    std::string str(const std::string& key, const std::string& dflt) const {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            while (!a.empty() && a[0] == '-') a.erase(0, 1);
            if (a == key && i + 1 < argc) return argv[i + 1];
            if (a.rfind(key + "=", 0) == 0) return a.substr(key.size() + 1);
        }
        return dflt;
    }
    long num(const std::string& key, long dflt) const {
        const std::string s = str(key, "");
        return s.empty() ? dflt : std::stol(s);
    }
    bool flag(const std::string& key) const {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            while (!a.empty() && a[0] == '-') a.erase(0, 1);
            if (a == key || a == key + "=true") return true;
        }
        return false;
    }
};

} // namespace kftools
