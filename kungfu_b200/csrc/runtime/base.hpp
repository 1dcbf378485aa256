// kungfu_b200 runtime — logging, configuration from the environment, stall watchdog, small utilities.
//
// Reference counterparts: srcs/go/log/logger.go:14-150 (levelled logger), srcs/go/kungfu/config/config.go:24-67
// (KUNGFU_CONFIG_* tunables), srcs/go/utils/stalldetector.go:15-46 (prints "<op> stalled for ..." every 3 s).
#pragma once
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace kfrt {

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
[[noreturn]] void fail(const std::string& what);
std::string strprintf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));

// ---------------------------------------------------------------------------------------------------
// logging
// ---------------------------------------------------------------------------------------------------
enum class LogLevel : int { Debug = 0, Info = 1, Warn = 2, Error = 3 };

class Logger {
public:
    static Logger& get();
    void set_level(LogLevel l) { level_ = l; }
    void set_level(const std::string& name);
    LogLevel level() const { return level_; }
    void set_output(const std::string& path); // append to file instead of stderr
    void set_prefix(const std::string& p) { prefix_ = p; }
    void logf(LogLevel l, const char* fmt, ...) __attribute__((format(printf, 3, 4)));

private:
    Logger();
    std::mutex mu_;
    LogLevel level_ = LogLevel::Info;
    FILE* out_ = nullptr;
    std::string prefix_;
    std::chrono::steady_clock::time_point t0_;
};

#define KF_LOGD(...) ::kfrt::Logger::get().logf(::kfrt::LogLevel::Debug, __VA_ARGS__)
#define KF_LOGI(...) ::kfrt::Logger::get().logf(::kfrt::LogLevel::Info, __VA_ARGS__)
#define KF_LOGW(...) ::kfrt::Logger::get().logf(::kfrt::LogLevel::Warn, __VA_ARGS__)
#define KF_LOGE(...) ::kfrt::Logger::get().logf(::kfrt::LogLevel::Error, __VA_ARGS__)

// ---------------------------------------------------------------------------------------------------
// configuration (read once from the environment)
// ---------------------------------------------------------------------------------------------------
struct Config {
    bool enable_monitoring = false;      // KUNGFU_CONFIG_ENABLE_MONITORING
    bool enable_stall_detection = false; // KUNGFU_CONFIG_ENABLE_STALL_DETECTION=true: every wait is watched, warnings every 3 s
    bool stall_detection_off = false;    // ...=false: nothing is watched. Unset: whole operations only, warnings every 30 s
    double monitoring_period_s = 1.0;    // KUNGFU_CONFIG_MONITORING_PERIOD
    std::string strategy_hash = "SIMPLE"; // KUNGFU_CONFIG_STRATEGY_HASH_METHOD (SIMPLE | NAME)
    double wait_runner_timeout_s = 300;  // KUNGFU_CONFIG_WAIT_RUNNER_TIMEOUT
    int conn_retry_count = 500;
    int conn_retry_period_ms = 200;
    bool use_unix_sock = true;
    size_t chunk_bytes = 1 << 20; // collective pipelining granularity (reference: session.go:292-295)
    static Config& get();
    static void reload();
};

std::string getenv_str(const char* name, const std::string& dflt = "");
bool getenv_bool(const char* name, bool dflt);
double parse_duration_s(const std::string& s, double dflt); // "10s", "500ms", "2m", "1.5"

// names of the env vars of the runner -> worker contract (reference: srcs/go/kungfu/env/envs.go:4-18)
namespace envkeys {
constexpr const char* kSelfSpec = "KUNGFU_SELF_SPEC";
constexpr const char* kInitPeers = "KUNGFU_INIT_PEERS";
constexpr const char* kInitRunners = "KUNGFU_INIT_RUNNERS";
constexpr const char* kParentID = "KUNGFU_PARENT_ID";
constexpr const char* kInitClusterVersion = "KUNGFU_INIT_CLUSTER_VERSION";
constexpr const char* kAllReduceStrategy = "KUNGFU_ALLREDUCE_STRATEGY";
constexpr const char* kConfigServer = "KUNGFU_CONFIG_SERVER";
constexpr const char* kJobStart = "KUNGFU_JOB_START_TIMESTAMP";
constexpr const char* kProcStart = "KUNGFU_PROC_START_TIMESTAMP";
constexpr const char* kAllowNvlink = "KUNGFU_ALLOW_NVLINK";
constexpr const char* kCudaVisibleDevices = "KUNGFU_CUDA_VISIBLE_DEVICES";
} // namespace envkeys

// ---------------------------------------------------------------------------------------------------
// stall watchdog
// ---------------------------------------------------------------------------------------------------
class StallGuard {
public:
    // `fine` marks the inner waits (one per chunk and peer); they are watched only when stall detection is switched on explicitly.
    // Whole operations (AllReduce(name), Barrier, Consensus, updateTo, request) are watched by default and reported after 30 s:
    // a job that hangs because two workers called different operations says so instead of staying silent.
    explicit StallGuard(const std::string& what, bool fine = false);
    ~StallGuard();

private:
    uint64_t id_ = 0;
};
// Called by the watchdog about twice per second with the age (seconds) of the oldest watched operation that is still pending, while
// there is one older than 0.5 s. The peer uses it to tell a hang watch "I am waiting inside a collective" (a victim, not the culprit).
void set_stall_reporter(std::function<void(double)> f);

// ---------------------------------------------------------------------------------------------------
// scope tracer (reference: TRACE_SCOPE / stdtracer, srcs/cpp/include/kungfu/utils/trace.hpp:2-16): per-scope call count
// and accumulated time, active when KUNGFU_ENABLE_TRACE is truthy, summary via trace_report() / at exit.
// ---------------------------------------------------------------------------------------------------
class TraceScope {
public:
    explicit TraceScope(const char* name);
    ~TraceScope();

private:
    const char* name_;
    double t0_;
};
std::string trace_report();
#define KF_TRACE_SCOPE(name) ::kfrt::TraceScope _kf_trace_scope_##__LINE__(name)

double now_s();            // monotonic seconds
uint64_t unix_time_ns();   // wall clock
void sleep_ms(int ms);

// merge error strings of parallel tasks (reference: utils.MergeErrors)
std::string merge_errors(const std::vector<std::string>& errs, const std::string& hint);

} // namespace kfrt
