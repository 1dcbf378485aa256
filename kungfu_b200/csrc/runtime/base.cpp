// Implementation of base.hpp (the header carries the design notes and the reference file:line mapping).
#include "base.hpp"

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace kfrt {

void fail(const std::string& what) { throw Error(what); }

std::string strprintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    va_list ap2;
    va_copy(ap2, ap);
    int n = vsnprintf(nullptr, 0, fmt, ap);
    va_end(ap);
    std::string s(n > 0 ? n : 0, '\0');
    if (n > 0) vsnprintf(&s[0], n + 1, fmt, ap2);
    va_end(ap2);
    return s;
}

// ---- logger ---------------------------------------------------------------------------------------
Logger::Logger() : t0_(std::chrono::steady_clock::now()) {
    set_level(getenv_str("KUNGFU_CONFIG_LOG_LEVEL", "INFO"));
}
Logger& Logger::get() {
    static Logger* l = new Logger(); // leaked on purpose: usable during static destruction
    return *l;
}
void Logger::set_level(const std::string& name) {
    std::string n;
    for (char c : name) n.push_back((char)toupper(c));
    if (n == "DEBUG") level_ = LogLevel::Debug;
    else if (n == "INFO") level_ = LogLevel::Info;
    else if (n == "WARN" || n == "WARNING") level_ = LogLevel::Warn;
    else if (n == "ERROR") level_ = LogLevel::Error;
}
void Logger::set_output(const std::string& path) {
    std::lock_guard<std::mutex> g(mu_);
    FILE* f = fopen(path.c_str(), "ae");
    if (f) {
        if (out_) fclose(out_);
        out_ = f;
    }
}
void Logger::logf(LogLevel l, const char* fmt, ...) {
    if ((int)l < (int)level_) return;
    static const char tags[] = {'D', 'I', 'W', 'E'};
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
    std::lock_guard<std::mutex> g(mu_);
    FILE* o = out_ ? out_ : stderr;
    fprintf(o, "[%c] %9.3fs %s%s\n", tags[(int)l], el, prefix_.c_str(), buf);
    fflush(o);
}

// ---- env ------------------------------------------------------------------------------------------
std::string getenv_str(const char* name, const std::string& dflt) {
    const char* v = getenv(name);
    return v ? std::string(v) : dflt;
}
bool getenv_bool(const char* name, bool dflt) {
    const char* v = getenv(name);
    if (!v) return dflt;
    std::string s(v);
    for (auto& c : s) c = (char)tolower(c);
    return s == "1" || s == "true" || s == "yes" || s == "on";
}
double parse_duration_s(const std::string& s, double dflt) {
    if (s.empty()) return dflt;
    char* end = nullptr;
    double v = strtod(s.c_str(), &end);
    if (end == s.c_str()) return dflt;
    std::string unit(end);
    if (unit == "ms") return v / 1e3;
    if (unit == "us") return v / 1e6;
    if (unit == "m") return v * 60;
    if (unit == "h") return v * 3600;
    return v; // "s" or bare number
}

static Config* g_cfg = nullptr;
static std::mutex g_cfg_mu;
static void load_config(Config& c) {
    c.enable_monitoring = getenv_bool("KUNGFU_CONFIG_ENABLE_MONITORING", false);
    c.enable_stall_detection = getenv_bool("KUNGFU_CONFIG_ENABLE_STALL_DETECTION", false);
    c.stall_detection_off = getenv("KUNGFU_CONFIG_ENABLE_STALL_DETECTION") != nullptr && !c.enable_stall_detection;
    c.monitoring_period_s = parse_duration_s(getenv_str("KUNGFU_CONFIG_MONITORING_PERIOD"), 1.0);
    c.strategy_hash = getenv_str("KUNGFU_CONFIG_STRATEGY_HASH_METHOD", "SIMPLE");
    c.wait_runner_timeout_s = parse_duration_s(getenv_str("KUNGFU_CONFIG_WAIT_RUNNER_TIMEOUT"), 300);
    c.use_unix_sock = getenv_bool("KUNGFU_CONFIG_USE_UNIX_SOCK", true);
    c.conn_retry_count = atoi(getenv_str("KUNGFU_CONFIG_CONN_RETRY_COUNT", "500").c_str());
    c.conn_retry_period_ms = (int)(parse_duration_s(getenv_str("KUNGFU_CONFIG_CONN_RETRY_PERIOD"), 0.2) * 1000);
    long cb = atol(getenv_str("KUNGFU_CONFIG_CHUNK_BYTES", "1048576").c_str());
    c.chunk_bytes = cb > 0 ? (size_t)cb : (1 << 20);
}
Config& Config::get() {
    std::lock_guard<std::mutex> g(g_cfg_mu);
    if (!g_cfg) {
        g_cfg = new Config();
        load_config(*g_cfg);
    }
    return *g_cfg;
}
void Config::reload() {
    std::lock_guard<std::mutex> g(g_cfg_mu);
    if (!g_cfg) g_cfg = new Config();
    load_config(*g_cfg);
}

// ---- time -----------------------------------------------------------------------------------------
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
uint64_t unix_time_ns() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}
void sleep_ms(int ms) { std::this_thread::sleep_for(std::chrono::milliseconds(ms)); }

// ---- stall watchdog -------------------------------------------------------------------------------
namespace {
struct Watchdog {
    std::mutex mu;
    std::condition_variable cv;
    std::map<uint64_t, std::pair<std::string, double>> ops;
    uint64_t next = 1;
    bool started = false;
    std::function<void(double)> reporter; // guarded by mu
    bool reporting = false;               // a reporter call is in flight (made without the lock)
    std::condition_variable idle;
    void ensure() {
        if (started) return;
        started = true;
        std::thread([this] {
            const bool full = Config::get().enable_stall_detection;
            const double threshold = full ? 3.0 : 30.0; // first report after this long, then once per `threshold` seconds
            std::map<uint64_t, double> last_report;
            std::unique_lock<std::mutex> lk(mu);
            while (true) {
                cv.wait_for(lk, reporter ? std::chrono::milliseconds(500) : std::chrono::milliseconds(3000));
                const double t = now_s();
                if (reporter && !ops.empty()) {
                    double oldest = 0;
                    for (auto& kv : ops) oldest = std::max(oldest, t - kv.second.second);
                    if (oldest >= 0.5) {
                        auto f = reporter;
                        reporting = true;
                        lk.unlock();
                        f(oldest);
                        lk.lock();
                        reporting = false;
                        idle.notify_all();
                    }
                }
                std::vector<std::string> msgs; // formatted under the lock, logged after it is released (ops may change meanwhile)
                for (auto& kv : ops) {
                    const double d = t - kv.second.second;
                    if (d < threshold) continue;
                    auto it = last_report.find(kv.first);
                    if (it != last_report.end() && t - it->second < threshold - 0.5) continue;
                    last_report[kv.first] = t;
                    msgs.push_back(strprintf("%s stalled for %.0fs", kv.second.first.c_str(), d));
                }
                for (auto it = last_report.begin(); it != last_report.end();) // forget finished operations
                    it = ops.count(it->first) ? std::next(it) : last_report.erase(it);
                lk.unlock();
                for (auto& m : msgs) KF_LOGW("%s", m.c_str());
                lk.lock();
            }
        }).detach();
    }
};
Watchdog& wd() {
    static Watchdog* w = new Watchdog();
    return *w;
}
} // namespace

void set_stall_reporter(std::function<void(double)> f) {
    Watchdog& w = wd();
    std::unique_lock<std::mutex> g(w.mu);
    w.reporter = std::move(f);
    if (w.reporter) w.ensure();
    else w.idle.wait(g, [&] { return !w.reporting; }); // whoever clears the reporter may destroy what it captured right after
    w.cv.notify_all();
}

StallGuard::StallGuard(const std::string& what, bool fine) {
    const Config& cfg = Config::get();
    if (cfg.stall_detection_off || (fine && !cfg.enable_stall_detection)) return;
    Watchdog& w = wd();
    std::lock_guard<std::mutex> g(w.mu);
    w.ensure();
    id_ = w.next++;
    w.ops[id_] = {what, now_s()};
}
StallGuard::~StallGuard() {
    if (!id_) return;
    Watchdog& w = wd();
    std::lock_guard<std::mutex> g(w.mu);
    auto it = w.ops.find(id_);
    if (it != w.ops.end()) {
        const double d = now_s() - it->second.second;
        if (d >= (Config::get().enable_stall_detection ? 3.0 : 30.0)) KF_LOGW("%s recovered after %.1fs", it->second.first.c_str(), d);
        w.ops.erase(it);
    }
}

// ---- scope tracer ---------------------------------------------------------------------------------
namespace {
struct Tracer {
    bool on = getenv_bool("KUNGFU_ENABLE_TRACE", false);
    std::mutex mu;
    std::map<std::string, std::pair<uint64_t, double>> scopes;
    ~Tracer() {
        if (on && !scopes.empty()) fprintf(stderr, "%s", report().c_str());
    }
    std::string report() {
        std::lock_guard<std::mutex> g(mu);
        std::string s = "[kungfu trace] scope                                   calls      total(ms)     mean(us)\n";
        for (auto& kv : scopes)
            s += strprintf("[kungfu trace] %-40s %8llu %14.3f %12.3f\n", kv.first.c_str(), (unsigned long long)kv.second.first,
                           kv.second.second * 1e3, kv.second.first ? kv.second.second * 1e6 / kv.second.first : 0.0);
        return s;
    }
};
Tracer& tracer() {
    static Tracer t;
    return t;
}
} // namespace
TraceScope::TraceScope(const char* name) : name_(name), t0_(tracer().on ? now_s() : 0) {}
TraceScope::~TraceScope() {
    Tracer& t = tracer();
    if (!t.on) return;
    const double d = now_s() - t0_;
    std::lock_guard<std::mutex> g(t.mu);
    auto& e = t.scopes[name_];
    e.first += 1;
    e.second += d;
}
std::string trace_report() { return tracer().report(); }

std::string merge_errors(const std::vector<std::string>& errs, const std::string& hint) {
    std::string msg;
    int n = 0;
    for (auto& e : errs)
        if (!e.empty()) {
            if (n++) msg += "; ";
            msg += e;
        }
    if (!n) return "";
    return hint + ": " + std::to_string(n) + " error(s): " + msg;
}

} // namespace kfrt
