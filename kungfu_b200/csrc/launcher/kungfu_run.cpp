// kungfu-run — the launcher of kungfu_b200 (C++; the reference's is Go).
//
// Reference: srcs/go/cmd/kungfu-run/app/kungfu-run.go:18-124 (main flow), srcs/go/kungfu/runner/{flags.go:69-104,
// simple.go:13-21, watch.go:23-150, monitored.go:18-75, handler.go:19-123, discovery.go:18-26},
// srcs/go/kungfu/runner/monitorserver/monitor.go:17-198, srcs/go/kungfu/job/{job.go:17-102, gpu_resource.go:11-51},
// srcs/go/utils/runner/local/local.go:18-92 (tee'd, colour-prefixed worker output; cancel all on first failure).
//
//   kungfu-run [-np N] [-H ip:slots,...|-hostfile f] [-self ip|-nic name] [-strategy S] [-port-range a-b] [-port p]
//              [-w] [-k] [-config-server URL] [-builtin-config-port p] [-init-version v] [-timeout d] [-logdir d]
//              [-logfile f] [-q] [-v] [-allow-nvlink] [-auto-recover d] [-delay d] [-debug-port p] prog args...
//
// Modes: static (spawn local workers, fail fast) | -w watch (elastic: a control server receives Stage updates from the
// workers' resize protocol and spawns/reaps local workers by diff, GPUs from a smallest-free pool) | -auto-recover
// (heartbeat failure detector on :7756; on a stalled batch: kill the local workers, patch --n-epochs/--num-epochs,
// append `--restart 1`, relaunch).
//
// B200 default: every worker sees ALL GPUs (P2P / NVLS need that) and is told its device through
// KUNGFU_CUDA_VISIBLE_DEVICES; `-allow-nvlink=false` restores the reference's one-visible-GPU isolation.
#include <arpa/inet.h>
#include <dirent.h>
#include <fcntl.h>
#include <ifaddrs.h>
#include <net/if.h>
#include <netinet/in.h>
#include <signal.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>

#include "base.hpp"
#include "config_server.hpp"
#include "graph.hpp"
#include "http.hpp"
#include "plan.hpp"
#include "transport.hpp"

#include "runner_core.hpp"

using namespace kfrt;
using namespace kfrun;

extern char** environ;

namespace {

// ---- running one process with tee'd, prefixed output ------------------------------------------------
const char* kColors[] = {"\x1b[1;32m", "\x1b[1;34m", "\x1b[1;33m", "\x1b[1;35m", "\x1b[1;36m", "\x1b[1;31m"};

struct Cancel {
    std::atomic<bool> flag{false};
    std::mutex mu;
    std::set<pid_t> pids;
    void add(pid_t p) {
        std::lock_guard<std::mutex> g(mu);
        pids.insert(p);
    }
    void remove(pid_t p) {
        std::lock_guard<std::mutex> g(mu);
        pids.erase(p);
    }
    void cancel(int sig = SIGTERM) {
        flag = true;
        std::lock_guard<std::mutex> g(mu);
        for (pid_t p : pids) ::kill(p, sig); // exact PIDs we forked
    }
};

void pump(int fd, const std::string& tag, const char* color, bool verbose, FILE* console, const std::string& logpath,
          std::shared_ptr<std::string> first_line = nullptr) {
    FILE* lf = logpath.empty() ? nullptr : fopen(logpath.c_str(), "we");
    FILE* in = fdopen(fd, "r");
    if (!in) {
        close(fd);
        if (lf) fclose(lf);
        return;
    }
    char* line = nullptr;
    size_t cap = 0;
    ssize_t n;
    static std::mutex console_mu;
    bool first = true;
    while ((n = getline(&line, &cap, in)) >= 0) {
        if (first && first_line) first_line->assign(line, (size_t)n); // read by the waiter after this thread is joined
        first = false;
        if (lf) {
            fwrite(line, 1, (size_t)n, lf);
            fflush(lf);
        }
        if (verbose) {
            std::lock_guard<std::mutex> g(console_mu);
            fprintf(console, "%s[%s]\x1b[m %s", color, tag.c_str(), line);
            if (n == 0 || line[n - 1] != '\n') fputc('\n', console);
            fflush(console);
        }
    }
    free(line);
    fclose(in);
    if (lf) fclose(lf);
}

// A started child with its output pumps. `stdin_fd` >= 0: the child reads its activation (environment) from a pipe (standby workers).
struct Child {
    pid_t pid = -1;
    int stdin_fd = -1;
    std::thread out, err;
    std::shared_ptr<std::string> first_err = std::make_shared<std::string>(); // first line the process wrote to stderr
};

// Start-up retry (runner_core.hpp `restart_after_startup_failure`): bounded by KUNGFU_START_RETRIES (default 10); the prefix can be
// changed with KUNGFU_RETRY_STDERR_PREFIX (empty = off).
bool should_restart(const Child& c, int rc, int attempt) {
    if (!c.first_err) return false;
    static const std::string prefix = getenv_str("KUNGFU_RETRY_STDERR_PREFIX", kfrun::loader_bug_prefix());
    static const int limit = std::atoi(getenv_str("KUNGFU_START_RETRIES", "10").c_str());
    return kfrun::restart_after_startup_failure(*c.first_err, rc, attempt, prefix, limit);
}

Child spawn_proc(const Proc& p, bool verbose, const std::string& logdir, Cancel* cancel, int color_idx, bool want_stdin) {
    // O_CLOEXEC: workers are spawned from several threads at once; a plain pipe() leaks this worker's write ends into every
    // sibling forked meanwhile, the pumps then never see EOF when this worker exits, its waiter thread never returns, and the
    // watch loop (which joins the waiter of a removed worker before it handles the next stage) wedges. dup2 onto 0 / 1 / 2 in
    // the child clears the flag on the copies that matter.
    int out_pipe[2], err_pipe[2], in_pipe[2] = {-1, -1};
    if (pipe2(out_pipe, O_CLOEXEC) != 0 || pipe2(err_pipe, O_CLOEXEC) != 0) fail("pipe() failed");
    if (want_stdin && pipe2(in_pipe, O_CLOEXEC) != 0) fail("pipe() failed");
    std::vector<std::string> envs;
    for (char** e = environ; *e; ++e) {
        std::string kv = *e;
        std::string k = kv.substr(0, kv.find('='));
        if (!p.envs.count(k)) envs.push_back(kv);
    }
    for (auto& kv : p.envs) envs.push_back(kv.first + "=" + kv.second);
    std::vector<char*> envp, argv;
    for (auto& s : envs) envp.push_back(const_cast<char*>(s.c_str()));
    envp.push_back(nullptr);
    argv.push_back(const_cast<char*>(p.prog.c_str()));
    for (auto& a : p.args) argv.push_back(const_cast<char*>(a.c_str()));
    argv.push_back(nullptr);
    pid_t pid = fork();
    if (pid < 0) fail("fork() failed");
    if (pid == 0) {
        if (want_stdin) dup2(in_pipe[0], 0);
        dup2(out_pipe[1], 1);
        dup2(err_pipe[1], 2);
        close(out_pipe[0]);
        close(out_pipe[1]);
        close(err_pipe[0]);
        close(err_pipe[1]);
        if (want_stdin) {
            close(in_pipe[0]);
            close(in_pipe[1]);
        }
        execvpe(p.prog.c_str(), argv.data(), envp.data());
        fprintf(stderr, "kungfu-run: cannot exec %s: %s\n", p.prog.c_str(), strerror(errno));
        _exit(127);
    }
    close(out_pipe[1]);
    close(err_pipe[1]);
    if (want_stdin) close(in_pipe[0]);
    if (cancel) cancel->add(pid);
    if (!logdir.empty()) mkdir(logdir.c_str(), 0755);
    const std::string base = logdir.empty() ? "" : logdir + "/" + p.log_prefix;
    const char* color = kColors[color_idx % 6];
    Child c;
    c.pid = pid;
    c.stdin_fd = want_stdin ? in_pipe[1] : -1;
    c.out = std::thread(pump, out_pipe[0], p.name + "::stdout", color, verbose, stdout, base.empty() ? "" : base + ".stdout.log", nullptr);
    c.err = std::thread(pump, err_pipe[0], p.name + "::stderr", color, verbose, stderr, base.empty() ? "" : base + ".stderr.log", c.first_err);
    return c;
}

// returns the exit code (128+signal when killed)
int wait_proc(Child& c, Cancel* cancel) {
    int status = 0;
    while (waitpid(c.pid, &status, 0) < 0 && errno == EINTR) {
    }
    if (cancel) cancel->remove(c.pid);
    if (c.stdin_fd >= 0) close(c.stdin_fd), c.stdin_fd = -1;
    c.out.join();
    c.err.join();
    if (WIFEXITED(status)) return WEXITSTATUS(status);
    if (WIFSIGNALED(status)) return 128 + WTERMSIG(status);
    return 1;
}

int run_proc(const Proc& p, bool verbose, const std::string& logdir, Cancel* cancel, int color_idx) {
    for (int attempt = 1;; ++attempt) {
        Child c = spawn_proc(p, verbose, logdir, cancel, color_idx, false);
        const int rc = wait_proc(c, cancel);
        if ((cancel && cancel->flag) || !should_restart(c, rc, attempt)) return rc;
        KF_LOGE("%s: restarting for the %d-th time (exit %d after the loader bug at start-up)", p.name.c_str(), attempt, rc);
    }
}

// run all procs concurrently; the first failure cancels the rest (fail-fast). Returns 0 iff all succeeded.
int run_all(const std::vector<Proc>& procs, bool verbose, const std::string& logdir, Cancel& cancel) {
    std::vector<int> codes(procs.size(), 0);
    std::vector<std::thread> ts;
    for (size_t i = 0; i < procs.size(); ++i)
        ts.emplace_back([&, i] {
            codes[i] = run_proc(procs[i], verbose, logdir, &cancel, (int)i);
            if (codes[i] != 0 && !cancel.flag) {
                KF_LOGE("%s exited with %d, cancelling the others", procs[i].name.c_str(), codes[i]);
                cancel.cancel();
            }
        });
    for (auto& t : ts) t.join();
    for (int c : codes)
        if (c != 0) return c;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// signal handling
// ---------------------------------------------------------------------------------------------------
Cancel* g_cancel = nullptr;
std::atomic<bool> g_stop{false};
std::atomic<bool> g_exit_by_peer{false}; // stop asked for by an `exit` control message (an operator's tool), not by a signal
void on_signal(int sig) {
    g_stop = true;
    if (g_cancel) {
        g_cancel->flag = true;
        // async-signal-safe enough: kill() on the pids we track (best effort without the lock)
        for (pid_t p : g_cancel->pids) ::kill(p, sig == SIGINT ? SIGINT : SIGTERM);
    }
}

void install_timeout(double timeout_s, Cancel& cancel) {
    if (timeout_s <= 0) return;
    std::thread([timeout_s, &cancel] {
        const double t0 = now_s();
        while (now_s() - t0 < timeout_s) {
            if (g_stop) return;
            sleep_ms(100);
        }
        KF_LOGE("timeout after %.1fs, killing workers", timeout_s);
        g_stop = true;
        cancel.cancel(SIGKILL);
    }).detach();
}

// ---------------------------------------------------------------------------------------------------
// static mode
// ---------------------------------------------------------------------------------------------------
int simple_run(const Flags& f, const Job& job, const Cluster& cluster, uint32_t self_ip) {
    Cancel cancel;
    g_cancel = &cancel;
    install_timeout(f.timeout_s, cancel);
    auto procs = job.create_procs(cluster, self_ip, f.init_version);
    KF_LOGI("will parallel run %d instances of %s", (int)procs.size(), job.prog.c_str());
    const double t0 = now_s();
    const int rc = run_all(procs, f.verbose, f.logdir, cancel);
    KF_LOGI("all %d/%d local peers finished, took %.2fs", (int)procs.size(), (int)cluster.workers.size(), now_s() - t0);
    return g_stop && rc == 0 ? 1 : rc;
}

// ---------------------------------------------------------------------------------------------------
// watch (elastic) mode
// ---------------------------------------------------------------------------------------------------
// Watch mode + -shrink-on-failure + -auto-recover d: workers send the same heartbeats as under the restart detector
// (`monitor_batch_begin/end`, key "begin:<rank>"); a rank whose batch has been open for longer than d is not dead, so nobody would ever
// see an exit status — the watch kills it, which turns "hung" into "crashed" and hands it to the shrink path.
class HangWatch {
public:
    using OnHang = std::function<void(int rank, double stalled_s)>;
    explicit HangWatch(double threshold_s) : threshold_s_(threshold_s) {}
    ~HangWatch() { stop(); }
    void start(const std::string& self_ip, int port, OnHang cb) {
        cb_ = std::move(cb);
        try {
            http_.start(self_ip.rfind("127.", 0) == 0 ? self_ip : std::string("0.0.0.0"), port, [this](const HttpRequest& r) { return handle(r); });
        } catch (const std::exception& e) {
            KF_LOGE("hang watch: %s", e.what());
            return;
        }
        running_ = true;
        th_ = std::thread([this] {
            while (running_) {
                sleep_ms(100);
                std::vector<std::pair<int, double>> late;
                {
                    std::lock_guard<std::mutex> g(mu_);
                    const double t = now_s();
                    for (size_t i = 0; i < begin_.size(); ++i)
                        if (begin_[i] != 0 && t - begin_[i] > threshold_s_ && (i >= waiting_.size() || t - waiting_[i] > 2.0)) {
                            late.emplace_back((int)i, t - begin_[i]);
                            begin_[i] = 0;
                        }
                }
                for (auto& l : late) cb_(l.first, l.second);
            }
        });
    }
    // the membership changed: ranks are renumbered, batches that were open belong to the old numbering
    void reset() {
        std::lock_guard<std::mutex> g(mu_);
        begin_.clear();
        waiting_.clear();
    }
    void stop() {
        if (!running_.exchange(false)) return;
        if (th_.joinable()) th_.join();
        http_.stop();
    }

private:
    HttpResponse handle(const HttpRequest& r) {
        HttpResponse resp;
        std::string key;
        try {
            JsonValue v = JsonValue::parse(r.body);
            const JsonValue* k = v.find("key");
            if (k) key = k->str;
        } catch (...) {
        }
        const auto colon = key.find(':');
        if (colon == std::string::npos) {
            resp.status = 400;
            return resp;
        }
        const std::string kind = key.substr(0, colon);
        const int v = atoi(key.c_str() + colon + 1);
        if (v < 0 || v > 65535) return resp;
        std::lock_guard<std::mutex> g(mu_);
        if ((size_t)v >= begin_.size()) begin_.resize(v + 1, 0.0);
        if ((size_t)v >= waiting_.size()) waiting_.resize(v + 1, 0.0);
        if (kind == "begin") begin_[v] = now_s();
        else if (kind == "end" || kind == "trainend") begin_[v] = 0;
        else if (kind == "waiting") waiting_[v] = now_s(); // inside a collective, waiting for somebody else: a victim
        return resp;
    }
    double threshold_s_;
    OnHang cb_;
    HttpServer http_;
    std::thread th_;
    std::mutex mu_;
    std::vector<double> begin_, waiting_;
    std::atomic<bool> running_{false};
};

struct Watcher {
    const Flags& f;
    Job job;
    PeerID self;
    Cancel cancel;
    GpuPool gpus;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Stage> stages;
    std::vector<Stage> history;
    std::map<uint64_t, std::thread> running; // peer key -> waiter thread
    std::map<uint64_t, bool> finished;
    int n_running = 0;
    bool failed = false;
    int fail_code = 0;
    int n_shrinks = 0;
    std::vector<std::thread> helpers; // loss announcements in flight
    std::map<uint64_t, pid_t> pids;   // peer key -> process of the worker that currently has this identity
    std::map<uint64_t, double> born;  // peer key -> when this runner created it (workers added after the first stage)
    int bad_refills = 0;              // replacements that died within 20 s of their creation: three of them end -refill
    std::unique_ptr<HangWatch> hang;
    Cluster current;
    // -standby N: warmed-up processes (interpreter, torch, CUDA context on a reserved GPU slot) waiting on a pipe for an identity
    struct Standby {
        Child child;
        int gpu = -1;
        int seq = 0;
    };
    std::deque<std::shared_ptr<Standby>> standbys;
    int standby_seq = 0;

    Watcher(const Flags& fl, const Job& j, const PeerID& s) : f(fl), job(j), self(s), gpus(j.hosts.slots_of(s.ipv4)) {}

    bool standby_supported() const {
        const std::string base = job.prog.substr(job.prog.find_last_of('/') == std::string::npos ? 0 : job.prog.find_last_of('/') + 1);
        return f.standby > 0 && base.rfind("python", 0) == 0 && !job.args.empty() && job.args[0].rfind("-", 0) != 0;
    }

    // keep up to f.standby warm processes around while GPU slots are free
    void replenish_standbys() {
        if (!standby_supported()) return;
        while (true) {
            {
                std::lock_guard<std::mutex> g(mu);
                if ((int)standbys.size() >= f.standby) return;
            }
            const int slots = job.hosts.slots_of(self.ipv4);
            const int gpu = gpus.get();
            if (gpu < 0 && slots > 0) return; // every slot of this host is taken by a worker or a standby
            Proc p;
            auto sb = std::make_shared<Standby>();
            sb->gpu = gpu;
            sb->seq = ++standby_seq;
            p.name = "standby-" + std::to_string(sb->seq);
            p.log_prefix = p.name;
            p.prog = job.prog;
            p.args = {"-m", "kungfu_b200.cmd.standby"};
            p.args.insert(p.args.end(), job.args.begin(), job.args.end());
            const std::string idx = Job::cuda_index(gpu);
            p.envs[envkeys::kCudaVisibleDevices] = idx;
            if (job.allow_nvlink) p.envs[envkeys::kAllowNvlink] = "1";
            else p.envs["CUDA_VISIBLE_DEVICES"] = idx;
            p.envs["PYTHONUNBUFFERED"] = "1";
            p.envs["KUNGFU_STANDBY"] = "1";
            sb->child = spawn_proc(p, true, job.logdir, &cancel, sb->seq, true);
            std::lock_guard<std::mutex> g(mu);
            standbys.push_back(sb);
        }
    }

    void stop_standbys() {
        std::deque<std::shared_ptr<Standby>> all;
        {
            std::lock_guard<std::mutex> g(mu);
            all.swap(standbys);
        }
        for (auto& sb : all) {
            if (sb->child.stdin_fd >= 0) close(sb->child.stdin_fd), sb->child.stdin_fd = -1; // EOF without activation: it exits
            wait_proc(sb->child, &cancel);
            gpus.put(sb->gpu);
        }
    }

    void push(const Stage& st) {
        std::lock_guard<std::mutex> g(mu);
        for (auto& h : history)
            if (h.version == st.version) {
                if (h.cluster != st.cluster) KF_LOGE("inconsistent update for version %d ignored", st.version);
                return; // duplicates arrive from every worker
            }
        history.push_back(st);
        stages.push_back(st);
        cv.notify_all();
    }

    void create(const PeerID& id, const Stage& st) {
        std::shared_ptr<Standby> sb;
        {
            std::lock_guard<std::mutex> g(mu);
            if (!standbys.empty()) {
                sb = standbys.front();
                standbys.pop_front();
            }
        }
        if (sb) {
            // a standby that died while waiting (a script that fails before it reaches the control plane, an OOM kill) must not be handed
            // out as a worker: it would "crash" the moment it is activated
            int status = 0;
            if (waitpid(sb->child.pid, &status, WNOHANG) == sb->child.pid) {
                KF_LOGW("standby-%d is dead (status %d): starting %s cold", sb->seq, status, id.str().c_str());
                if (sb->child.stdin_fd >= 0) close(sb->child.stdin_fd), sb->child.stdin_fd = -1;
                cancel.remove(sb->child.pid);
                sb->child.out.join();
                sb->child.err.join();
                gpus.put(sb->gpu);
                sb.reset();
            }
        }
        const int gpu = sb ? sb->gpu : gpus.get();
        if (gpu < 0) KF_LOGW("no free GPU slot for %s", id.str().c_str());
        Proc p = job.new_proc(id, gpu, st.version, st.cluster);
        if (sb) {
            // activate the warm process: its identity is the environment a cold worker would have been exec'ed with
            std::string msg;
            for (auto& kv : p.envs) msg += kv.first + "=" + kv.second + "\n";
            msg += "\n";
            size_t off = 0;
            while (off < msg.size()) {
                const ssize_t n = write(sb->child.stdin_fd, msg.data() + off, msg.size() - off);
                if (n <= 0 && errno != EINTR) break;
                if (n > 0) off += (size_t)n;
            }
            close(sb->child.stdin_fd);
            sb->child.stdin_fd = -1;
            fprintf(stderr, "[kungfu-run] standby-%d becomes %s\n", sb->seq, p.name.c_str());
        }
        {
            std::lock_guard<std::mutex> g(mu);
            ++n_running;
            finished[id.key()] = false;
            if (history.size() > 1) born[id.key()] = now_s(); // not one of the workers the job started with
        }
        const int color = (int)(id.port % 6);
        std::thread t([this, p, gpu, id, color, sb] {
            Child own;
            if (!sb) own = spawn_proc(p, true, job.logdir, &cancel, color, false);
            Child& c = sb ? sb->child : own;
            {
                std::lock_guard<std::mutex> g(mu);
                pids[id.key()] = c.pid;
            }
            int rc = wait_proc(c, &cancel);
            Child* cur = &c;
            for (int attempt = 1; !cancel.flag && should_restart(*cur, rc, attempt); ++attempt) { // died in the loader before joining: start it again
                KF_LOGE("%s: restarting for the %d-th time (exit %d after the loader bug at start-up)", p.name.c_str(), attempt, rc);
                own = spawn_proc(p, true, job.logdir, &cancel, color, false);
                cur = &own;
                {
                    std::lock_guard<std::mutex> g(mu);
                    pids[id.key()] = own.pid;
                }
                rc = wait_proc(own, &cancel);
            }
            gpus.put(gpu);
            std::lock_guard<std::mutex> g(mu);
            pids.erase(id.key());
            --n_running;
            finished[id.key()] = true;
            if (rc != 0 && !cancel.flag) {
                KF_LOGE("%s finished with error %d", p.name.c_str(), rc);
                if (!shrink_without_locked(id)) {
                    failed = true;
                    fail_code = rc;
                }
            }
            cv.notify_all();
        });
        std::lock_guard<std::mutex> g(mu);
        running[id.key()] = std::move(t);
    }

    // -shrink-on-failure: worker `id` of this host died. Instead of taking the job down, announce the cluster without it as
    // the next stage: survivors get a "lost" control message (their collectives in flight abort, `recover()` adopts the stage),
    // the other runners get the usual "update", the config server is told so that later resizes start from the truth.
    // Called with `mu` held by the waiter thread of the dead process; the sends happen on a helper thread.
    bool shrink_without_locked(const PeerID& id) {
        if (!f.shrink_on_failure || history.empty()) return false;
        const Stage& last = history.back();
        if (!last.cluster.workers.contains(id)) return true; // already on its way out (a detached worker that crashed at exit)
        if (last.cluster.workers.size() <= 1) return false;  // nobody left to continue
        Stage st;
        st.version = last.version + 1;
        st.cluster = last.cluster;
        st.cluster.workers = last.cluster.workers.minus(PeerList{id});
        auto bt = born.find(id.key());
        if (bt != born.end() && now_s() - bt->second < 20.0 && ++bad_refills == 3)
            KF_LOGW("three replacement workers in a row died right after their start: -refill is off for the rest of the job");
        KF_LOGW("%s is gone: shrinking to %d workers (v%d)", id.str().c_str(), (int)st.cluster.workers.size(), st.version);
        fprintf(stderr, "[kungfu-run] worker %s lost, continuing with %d workers\n", id.str().c_str(), (int)st.cluster.workers.size());
        history.push_back(st);
        stages.push_back(st);
        ++n_shrinks;
        helpers.emplace_back([this, st] { announce_loss(st); });
        return true;
    }

    // HangWatch callback: `rank` of the current cluster has had a batch open for `stalled_s`
    void kill_hung(int rank, double stalled_s) {
        pid_t pid = -1;
        PeerID id{};
        {
            std::lock_guard<std::mutex> g(mu);
            if (rank < 0 || rank >= (int)current.workers.size()) return;
            id = current.workers[rank];
            auto it = pids.find(id.key());
            if (it == pids.end()) return; // not one of ours (another host's runner watches it)
            pid = it->second;
        }
        KF_LOGW("rank %d (%s) made no progress for %.1f s: killing it", rank, id.str().c_str(), stalled_s);
        fprintf(stderr, "[kungfu-run] worker %s hung for %.1f s, killing it\n", id.str().c_str(), stalled_s);
        kill(pid, SIGKILL);
    }

    void announce_loss(const Stage& st) {
        const std::string payload = st.to_json();
        if (!job.config_server.empty()) {
            std::string resp;
            const int code = http_request("PUT", job.config_server, st.cluster.to_json(), &resp);
            if (code != 200) KF_LOGW("config server did not take the shrunk cluster: %d %s", code, resp.c_str());
        }
        Client client(self);
        client.set_max_dial_tries(5); // a worker that does not answer within a second is the next casualty, not somebody to wait 100 s for
        for (auto& w : st.cluster.workers) {
            try {
                client.send(w, "lost", payload.data(), payload.size(), kConnControl, kNoFlag);
            } catch (const std::exception& e) {
                KF_LOGW("could not tell %s about the lost peer: %s", w.str().c_str(), e.what());
            }
        }
        for (auto& r : st.cluster.runners) {
            if (r == self) continue;
            try {
                client.send(r, "update", payload.data(), payload.size(), kConnControl, kNoFlag);
            } catch (const std::exception& e) {
                KF_LOGW("could not tell runner %s about the lost peer: %s", r.str().c_str(), e.what());
            }
        }
        client.close_all();
        bool refill_on;
        {
            std::lock_guard<std::mutex> g(mu);
            refill_on = f.refill && bad_refills < 3;
        }
        if (refill_on && !job.config_server.empty()) {
            // self-healing: once the survivors have had time to regroup, ask for the old size again. Nothing is forced: the bigger
            // cluster takes effect when the workers next consult the config server (their resize protocol), and the new worker is
            // created by whichever runner owns the address — from its standby pool when it has one.
            sleep_ms(1000);
            if (cancel.flag || g_stop) return;
            const Cluster grown = st.cluster.resized((int)st.cluster.workers.size() + 1);
            std::string resp;
            const int code = http_request("PUT", job.config_server, grown.to_json(), &resp);
            if (code == 200) {
                KF_LOGI("refill: proposed %d workers again (%s)", (int)grown.workers.size(), grown.workers.str().c_str());
                fprintf(stderr, "[kungfu-run] refill: proposed %d workers again\n", (int)grown.workers.size());
            } else {
                KF_LOGW("refill: config server answered %d %s", code, resp.c_str());
            }
        }
    }

    void remove(const PeerID& id) {
        std::thread t;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = running.find(id.key());
            if (it == running.end()) return;
            t = std::move(it->second);
            running.erase(it);
        }
        if (t.joinable()) t.join(); // detached workers stop themselves
    }

    void update(const Stage& st, Server& server) {
        server.set_token((uint32_t)st.version);
        if (!current.workers.empty() && current.workers.disjoint(st.cluster.workers)) KF_LOGE("full update detected");
        PeerList del = current.workers.minus(st.cluster.workers).on_host(self.ipv4);
        PeerList add = st.cluster.workers.minus(current.workers).on_host(self.ipv4);
        KF_LOGI("arrived at v%d, new np=%d, local: +%d/-%d", st.version, (int)st.cluster.workers.size(), (int)add.size(), (int)del.size());
        for (auto& id : del) remove(id);
        for (auto& id : add) create(id, st);
        {
            std::lock_guard<std::mutex> g(mu);
            current = st.cluster;
        }
        if (hang) hang->reset();
        replenish_standbys();
    }

    int loop(Server& server) {
        bool started = false;
        while (true) {
            Stage st;
            bool have = false;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait_for(lk, std::chrono::milliseconds(200), [&] { return !stages.empty() || failed || g_stop.load(); });
                if (failed) {
                    lk.unlock();
                    cancel.cancel();
                    join_all();
                    return fail_code ? fail_code : 1;
                }
                if (g_stop) {
                    // a launcher that idles in keep mode with no worker left and is told to leave by an `exit` message ends cleanly;
                    // a signal, or an exit that had to cancel running workers, is a failure of the job
                    const bool clean = g_exit_by_peer.load() && started && n_running == 0;
                    lk.unlock();
                    cancel.cancel();
                    join_all();
                    return clean ? 0 : 1;
                }
                if (!stages.empty()) {
                    st = stages.front();
                    stages.pop_front();
                    have = true;
                } else if (started && n_running == 0 && !f.keep) {
                    lk.unlock();
                    join_all();
                    return 0;
                }
            }
            if (have) {
                update(st, server);
                started = true;
            }
        }
    }

    void join_all() {
        if (hang) hang->stop();
        stop_standbys();
        std::vector<std::thread> hs;
        {
            std::lock_guard<std::mutex> g(mu);
            hs.swap(helpers);
        }
        for (auto& h : hs)
            if (h.joinable()) h.join();
        std::map<uint64_t, std::thread> ts;
        {
            std::lock_guard<std::mutex> g(mu);
            ts.swap(running);
        }
        for (auto& kv : ts)
            if (kv.second.joinable()) kv.second.join();
    }
};

int watch_run(const Flags& f, const Job& job, const Cluster& init, const PeerID& self) {
    Watcher w(f, job, self);
    g_cancel = &w.cancel;
    install_timeout(f.timeout_s, w.cancel);
    Server server(self, nullptr, nullptr, Config::get().use_unix_sock);
    server.set_control_handler([&](const PeerID& src, const std::string& name, const Bytes& payload) {
        if (name == "update") {
            try {
                Stage st = Stage::from_json(std::string(payload.begin(), payload.end()));
                KF_LOGD("stage v%d (%d workers) proposed by %s", st.version, (int)st.cluster.workers.size(), src.str().c_str());
                w.push(st);
            } catch (const std::exception& e) {
                KF_LOGE("bad stage from %s: %s", src.str().c_str(), e.what());
            }
        } else if (name == "exit") {
            KF_LOGW("exit requested by %s", src.str().c_str());
            g_exit_by_peer = true;
            g_stop = true;
        }
    });
    server.start();
    HttpServer debug;
    if (f.debug_port > 0) {
        debug.start("0.0.0.0", f.debug_port, [&](const HttpRequest&) {
            HttpResponse r;
            r.content_type = "application/json";
            std::lock_guard<std::mutex> g(w.mu);
            r.body = "[";
            for (size_t i = 0; i < w.history.size(); ++i) r.body += (i ? "," : "") + w.history[i].to_json();
            r.body += "]\n";
            return r;
        });
        KF_LOGI("debug server: http://127.0.0.1:%d/", f.debug_port);
    }
    if (f.shrink_on_failure && f.auto_recover_s > 0) {
        const int mport = atoi(getenv_str("KUNGFU_MONITOR_PORT", "7756").c_str());
        w.job.hang_watch = "http://" + format_ipv4(self.ipv4) + ":" + std::to_string(mport) + "/";
        w.hang.reset(new HangWatch(f.auto_recover_s));
        w.hang->start(format_ipv4(self.ipv4), mport,
                      [&w](int rank, double stalled_s) { w.kill_hung(rank, stalled_s); });
        KF_LOGI("hang watch: a worker whose batch stays open for %.1f s is killed and the cluster shrinks", f.auto_recover_s);
    }
    if (f.init_version < 0) {
        KF_LOGI("waiting to be initialized");
    } else {
        Stage st;
        st.version = f.init_version;
        st.cluster = init;
        w.push(st);
    }
    KF_LOGI("watching config server");
    const int rc = w.loop(server);
    KF_LOGI("stop watching");
    server.stop();
    debug.stop();
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// failure detector + auto recovery
// ---------------------------------------------------------------------------------------------------
struct MonitorResult {
    bool down = false, finish = false;
    int epochs = 0;
};

class MonitorServer {
public:
    MonitorServer(int cluster_size, double threshold_s) : n_(cluster_size), threshold_s_(threshold_s > 0 ? threshold_s : 10.0) {
        trainend_.assign(n_, 0);
        begin_.assign(n_, 0.0);
        epochs_.assign(n_, 0);
    }
    // blocks until a worker stalls (down) or every rank reported train end (finish)
    MonitorResult run(const std::string& self_ip, const HostList& hosts, int port = 7756) {
        bool main_server = false;
        std::vector<std::string> others;
        if (hosts.size() != 1) {
            int idx = 0;
            for (auto& h : hosts) {
                if (format_ipv4(h.ipv4) == self_ip) {
                    if (idx == 0) main_server = true;
                } else {
                    others.push_back(format_ipv4(h.ipv4));
                }
                ++idx;
            }
        }
        HttpServer http;
        try {
            // a loopback-alias host (127.0.0.x) listens on its own address so that several "hosts" can share a machine in tests
            http.start(self_ip.rfind("127.", 0) == 0 ? self_ip : std::string("0.0.0.0"), port, [this](const HttpRequest& r) { return handle(r); });
        } catch (const std::exception& e) {
            KF_LOGE("monitor server: %s", e.what());
        }
        MonitorResult res;
        auto relay = [&](const std::string& key) {
            for (auto& ip : others) http_request("POST", "http://" + ip + ":" + std::to_string(port) + "/", "{\"key\":\"" + key + "\"}", nullptr);
        };
        while (!g_stop && !stop_) {
            sleep_ms(100);
            std::lock_guard<std::mutex> g(mu_);
            int ended = 0;
            const double t = now_s();
            bool down = false;
            for (int i = 0; i < n_; ++i) {
                if (trainend_[i]) ++ended;
                if (begin_[i] != 0 && t - begin_[i] > threshold_s_) down = true;
            }
            if (down) {
                res.down = true;
                res.epochs = *std::min_element(epochs_.begin(), epochs_.end());
                if (main_server) relay("otherdown:" + std::to_string(res.epochs));
                break;
            }
            if (other_down_) {
                res.down = true;
                res.epochs = other_epochs_;
                break;
            }
            if (ended == n_ || other_finish_) {
                if (main_server) relay("otherfinish:0");
                res.finish = true;
                break;
            }
        }
        http.stop();
        return res;
    }
    void stop() { stop_ = true; }

private:
    HttpResponse handle(const HttpRequest& r) {
        HttpResponse resp;
        std::string key;
        try {
            JsonValue v = JsonValue::parse(r.body);
            const JsonValue* k = v.find("key");
            if (k) key = k->str;
        } catch (...) {
        }
        auto colon = key.find(':');
        if (colon == std::string::npos) {
            resp.status = 400;
            return resp;
        }
        const std::string kind = key.substr(0, colon);
        const int v = atoi(key.c_str() + colon + 1);
        std::lock_guard<std::mutex> g(mu_);
        const bool in_range = v >= 0 && v < n_;
        if (kind == "trainend" && in_range) trainend_[v] = 1;
        else if (kind == "begin" && in_range) begin_[v] = now_s();
        else if (kind == "end" && in_range) begin_[v] = 0;
        else if (kind == "epoch" && in_range) epochs_[v] += 1;
        else if (kind == "otherfinish") other_finish_ = true;
        else if (kind == "otherdown") {
            other_epochs_ = v;
            other_down_ = true;
        }
        return resp;
    }
    int n_;
    double threshold_s_;
    std::mutex mu_;
    std::vector<int> trainend_, epochs_;
    std::vector<double> begin_;
    bool other_finish_ = false, other_down_ = false;
    int other_epochs_ = 0;
    std::atomic<bool> stop_{false};
};

int monitored_run(const Flags& f, Job job, const Cluster& cluster, uint32_t self_ip) {
    const std::string self = f.self.empty() ? format_ipv4(self_ip) : f.self;
    int restarts = 0;
    while (!g_stop) {
        Cancel cancel;
        g_cancel = &cancel;
        auto procs = job.create_procs(cluster, self_ip, f.init_version);
        MonitorServer mon((int)cluster.workers.size(), f.auto_recover_s);
        KF_LOGI("will parallel run %d instances of %s under the failure detector (threshold %.1fs)", (int)procs.size(), job.prog.c_str(),
                f.auto_recover_s);
        std::atomic<int> rc{-1};
        std::thread runner([&] {
            rc = run_all(procs, f.verbose, f.logdir, cancel);
            if (rc != 0) mon.stop(); // a crashed worker ends the watch as well
        });
        MonitorResult res = mon.run(self, job.hosts, atoi(getenv_str("KUNGFU_MONITOR_PORT", "7756").c_str()));
        if (res.down || (rc.load() > 0 && !res.finish && !g_stop)) {
            KF_LOGW("worker failure detected (finished epochs: %d); restarting from the last checkpoint", res.epochs);
            cancel.cancel(SIGKILL);
            runner.join();
            for (size_t i = 0; i + 1 < job.args.size(); ++i)
                if (job.args[i] == "--n-epochs" || job.args[i] == "--num-epochs") {
                    const int ini = atoi(job.args[i + 1].c_str());
                    job.args[i + 1] = std::to_string(std::max(0, ini - res.epochs));
                }
            if (std::find(job.args.begin(), job.args.end(), "--restart") == job.args.end()) {
                job.args.push_back("--restart");
                job.args.push_back("1");
            }
            if (++restarts > 16) {
                KF_LOGE("giving up after %d restarts", restarts);
                return 1;
            }
            continue;
        }
        runner.join();
        if (res.finish && rc == 0) {
            KF_LOGI("success finish");
            return 0;
        }
        return rc.load() ? rc.load() : 1;
    }
    return 1;
}

// GPU model names from the driver's procfs (no CUDA context in the launcher); reference: utils.ListNvidiaGPUNames
std::vector<std::string> list_nvidia_gpu_names() {
    std::vector<std::string> names;
    DIR* d = opendir("/proc/driver/nvidia/gpus");
    if (!d) return names;
    while (dirent* e = readdir(d)) {
        if (e->d_name[0] == '.') continue;
        std::ifstream in(std::string("/proc/driver/nvidia/gpus/") + e->d_name + "/information");
        std::string line;
        while (std::getline(in, line))
            if (line.rfind("Model:", 0) == 0) {
                auto pos = line.find_first_not_of(" \t", 6);
                names.push_back(std::string(e->d_name) + " " + (pos == std::string::npos ? "" : line.substr(pos)));
            }
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    return names;
}

// What the job will see: framework / CUDA / NCCL environment, IPv4 interfaces, GPUs (reference: utils.LogKungfuEnv,
// LogCudaEnv, LogNCCLEnv, LogNICInfo). Debug level only.
void log_env() {
    for (char** e = environ; *e; ++e)
        if (!strncmp(*e, "KUNGFU_", 7) || !strncmp(*e, "CUDA_", 5) || !strncmp(*e, "NCCL_", 5)) KF_LOGD("env %s", *e);
    ifaddrs* ifa = nullptr;
    if (getifaddrs(&ifa) == 0) {
        for (auto* p = ifa; p; p = p->ifa_next)
            if (p->ifa_addr && p->ifa_addr->sa_family == AF_INET) {
                char buf[INET_ADDRSTRLEN] = {0};
                inet_ntop(AF_INET, &reinterpret_cast<sockaddr_in*>(p->ifa_addr)->sin_addr, buf, sizeof buf);
                KF_LOGD("nic %s %s", p->ifa_name, buf);
            }
        freeifaddrs(ifa);
    }
    for (const auto& g : list_nvidia_gpu_names()) KF_LOGD("gpu %s", g.c_str());
}

} // namespace

int main(int argc, char** argv) {
    try {
        Flags f;
        try {
            f = parse_flags(argc, argv);
        } catch (const UsageError& e) {
            if (*e.what()) fprintf(stderr, "kungfu-run: %s\n", e.what());
            fputs(usage_text().c_str(), stderr);
            return *e.what() ? 2 : 0;
        }
        if (f.quiet) Logger::get().set_level(LogLevel::Warn);
        Logger::get().set_prefix("[kungfu-run] ");
        if (!f.logfile.empty()) {
            std::string lf = f.logdir.empty() ? f.logfile : f.logdir + "/" + f.logfile;
            auto slash = lf.rfind('/');
            if (slash != std::string::npos) mkdir(lf.substr(0, slash).c_str(), 0755);
            Logger::get().set_output(lf);
        }
        log_env();
        if (f.delay_s > 0) {
            KF_LOGW("delay start for %.1fs", f.delay_s);
            sleep_ms((int)(f.delay_s * 1000));
        }
        signal(SIGPIPE, SIG_IGN);
        struct sigaction sa {};
        sa.sa_handler = on_signal;
        sigaction(SIGINT, &sa, nullptr);
        sigaction(SIGTERM, &sa, nullptr);

        HostList hosts;
        if (!f.hostfile.empty()) {
            std::ifstream in(f.hostfile);
            if (!in) fail("cannot read hostfile " + f.hostfile);
            std::stringstream ss;
            ss << in.rdbuf();
            hosts = HostList::parse_hostfile(ss.str());
        } else {
            hosts = HostList::parse(f.hosts);
        }
        const uint32_t self_ip = infer_self_ipv4(f.self, f.nic);
        const PeerID self{self_ip, (uint16_t)f.port};
        PeerList runners = hosts.gen_runner_list((uint16_t)f.port);
        if (!runners.contains(self)) fail(self.str() + " not in " + runners.str());
        PeerList peers = hosts.gen_peer_list(f.np, f.port_range);
        Cluster init{runners, peers};

        ConfigServer builtin;
        std::string config_server = f.config_server;
        if (f.builtin_config_port > 0) {
            builtin.start("0.0.0.0", f.builtin_config_port, &init);
            if (config_server.empty()) config_server = "http://127.0.0.1:" + std::to_string(f.builtin_config_port) + "/config";
            KF_LOGI("builtin config server at %s", config_server.c_str());
        }

        Job job;
        job.start_time = f.t0;
        job.strategy = f.strategy;
        job.parent = self;
        job.hosts = hosts;
        job.port_range = f.port_range;
        job.prog = f.prog;
        job.args = f.args;
        job.logdir = f.logdir;
        job.allow_nvlink = f.allow_nvlink;
        job.config_server = config_server;

        const double t0 = now_s();
        int rc;
        if (f.watch) rc = watch_run(f, job, init, self);
        else if (f.auto_recover_s > 0) rc = monitored_run(f, job, init, self_ip);
        else rc = simple_run(f, job, init, self_ip);
        if (f.builtin_config_port > 0) builtin.stop();
        KF_LOGD("kungfu-run finished, took %.2fs", now_s() - t0);
        return rc;
    } catch (const std::exception& e) {
        fprintf(stderr, "kungfu-run: %s\n", e.what());
        return 1;
    }
}
