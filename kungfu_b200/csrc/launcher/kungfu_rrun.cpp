// kungfu-rrun / kungfu-distribute — native SSH fan-out of `kungfu-run` (or any command) to every host of a job.
//
// Reference: srcs/go/cmd/kungfu-rrun, srcs/go/cmd/kungfu-distribute and srcs/go/utils/runner/remote/remote.go:22-131
// (`RemoteRunAll`: one SSH session per host, output redirected with a per-host prefix and into per-host log files, first failure
// cancels the rest; `RunStaticKungFuJob` / `RunElasticKungFuJob` build the per-host `kungfu-run` command lines). Here the SSH
// client is the system `ssh` binary driven through fork/exec + pipes + poll; there is no SSH library to link.
//
//   kungfu-rrun -np 16 -H 10.0.0.1:8,10.0.0.2:8 [-u user] [-w [-k] -config-server URL | -builtin-config-port P] [-strategy S] [-port P]
//               [-port-range a-b] [-nic N] [-logdir DIR] [-q] [-standby N] [-shrink-on-failure] [-refill] [-auto-recover D] [-timeout D]
//               [-local] [-dry-run] prog args...
//   kungfu-rrun -distribute -H 10.0.0.1,10.0.0.2 [-u user] cmd args...      (also: invoked as `kungfu-distribute`)
//
// `-local` starts the per-host launchers as local processes (hosts = loopback aliases 127.0.0.x on this machine): the multi-host
// tier without a cluster, and how the test-suite drives this binary.
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <signal.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "plan.hpp"

using namespace kfrt;

namespace {

struct Task {
    std::string name;              // host label
    std::vector<std::string> argv; // what to exec
    pid_t pid = -1;
    int fd = -1;                   // merged stdout+stderr of the child
    std::string partial;           // bytes of an unfinished line
    FILE* log = nullptr;
    int status = -1;
    bool done = false;
};

std::string shell_quote(const std::string& s) {
    if (!s.empty() && s.find_first_not_of("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_@%+=:,./-") == std::string::npos) return s;
    std::string q = "'";
    for (char c : s) {
        if (c == '\'') q += "'\\''";
        else q += c;
    }
    return q + "'";
}

std::string join_quoted(const std::vector<std::string>& v) {
    std::string s;
    for (auto& x : v) {
        if (!s.empty()) s += ' ';
        s += shell_quote(x);
    }
    return s;
}

const char* kColors[] = {"\033[32m", "\033[34m", "\033[33m", "\033[35m", "\033[36m", "\033[91m", "\033[92m", "\033[94m"};

void emit(Task& t, size_t idx, const std::string& line, bool color) {
    if (color) std::printf("%s[%s]\033[0m %s\n", kColors[idx % 8], t.name.c_str(), line.c_str());
    else std::printf("[%s] %s\n", t.name.c_str(), line.c_str());
    if (t.log) std::fprintf(t.log, "%s\n", line.c_str());
}

bool spawn(Task& t) {
    int p[2];
    if (pipe(p) != 0) return false;
    const pid_t pid = fork();
    if (pid < 0) return false;
    if (pid == 0) {
        setpgid(0, 0); // own process group: cancellation reaches everything the child started
        dup2(p[1], 1);
        dup2(p[1], 2);
        close(p[0]);
        close(p[1]);
        std::vector<char*> av;
        for (auto& a : t.argv) av.push_back(const_cast<char*>(a.c_str()));
        av.push_back(nullptr);
        execvp(av[0], av.data());
        std::fprintf(stderr, "exec %s: %s\n", av[0], strerror(errno));
        _exit(127);
    }
    close(p[1]);
    fcntl(p[0], F_SETFD, FD_CLOEXEC);
    t.pid = pid;
    t.fd = p[0];
    return true;
}

void usage() {
    std::fprintf(stderr,
                 "usage: kungfu-rrun -np N -H ip:slots,... [-u user] [-w] [-k] [-config-server URL] [-strategy S] [-port P] [-port-range a-b]\n"
                 "                   [-nic N] [-logdir DIR] [-q] [-local] [-dry-run] prog args...\n"
                 "       kungfu-rrun -distribute -H ip,... [-u user] [-logdir DIR] [-dry-run] cmd args...\n");
}

} // namespace

int main(int argc, char** argv) {
    std::string self = argv[0];
    bool distribute = self.size() >= 17 && self.substr(self.size() - 17) == "kungfu-distribute";
    int np = 1;
    std::string hosts, user, strategy, port, port_range, nic, logdir, config_server, runner = "kungfu-run";
    bool watch = false, keep = false, quiet = false, local = false, dry = false;
    std::vector<std::string> extra;
    std::string builtin_port;
    int i = 1;
    auto need = [&](const char* flag) -> std::string {
        if (i + 1 >= argc) {
            std::fprintf(stderr, "kungfu-rrun: %s needs a value\n", flag);
            exit(2);
        }
        return argv[++i];
    };
    for (; i < argc; ++i) {
        std::string a = argv[i];
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') a = a.substr(1); // accept --flag as well
        if (a == "-np") np = atoi(need("-np").c_str());
        else if (a == "-H") hosts = need("-H");
        else if (a == "-u") user = need("-u");
        else if (a == "-strategy") strategy = need("-strategy");
        else if (a == "-port") port = need("-port");
        else if (a == "-port-range") port_range = need("-port-range");
        else if (a == "-nic") nic = need("-nic");
        else if (a == "-logdir") logdir = need("-logdir");
        else if (a == "-config-server") config_server = need("-config-server");
        else if (a == "-kungfu-run") runner = need("-kungfu-run");
        else if (a == "-standby" || a == "-auto-recover" || a == "-timeout" || a == "-init-version" || a == "-delay") {
            extra.push_back(a); // kungfu-run flags with a value that every host's launcher gets unchanged
            extra.push_back(need(a.c_str()));
        } else if (a == "-shrink-on-failure" || a == "-refill" || a == "-allow-nvlink" || a == "-v") extra.push_back(a);
        else if (a == "-builtin-config-port") builtin_port = need("-builtin-config-port"); // first host only: it runs the config server
        else if (a == "-w") watch = true;
        else if (a == "-k") keep = true;
        else if (a == "-q") quiet = true;
        else if (a == "-local") local = true;
        else if (a == "-dry-run") dry = true;
        else if (a == "-distribute") distribute = true;
        else if (a == "-h" || a == "-help") {
            usage();
            return 0;
        } else break;
    }
    std::vector<std::string> prog(argv + i, argv + argc);
    if (hosts.empty() || prog.empty()) {
        usage();
        return 2;
    }
    HostList hl;
    try {
        hl = HostList::parse(hosts);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "kungfu-rrun: bad -H: %s\n", e.what());
        return 2;
    }
    if (!distribute && np > hl.cap()) {
        std::fprintf(stderr, "kungfu-rrun: -np %d exceeds the %d slots of -H %s\n", np, hl.cap(), hosts.c_str());
        return 2;
    }
    std::vector<Task> tasks;
    for (auto& h : hl) {
        const std::string ip = format_ipv4(h.ipv4);
        const std::string target = h.public_addr.empty() ? ip : h.public_addr;
        std::vector<std::string> inner;
        if (distribute) {
            inner = prog;
        } else {
            inner = {runner, "-self", ip, "-np", std::to_string(np), "-H", hosts};
            if (watch) inner.push_back("-w");
            if (keep) inner.push_back("-k");
            inner.insert(inner.end(), extra.begin(), extra.end());
            if (!builtin_port.empty()) {
                // the first host serves the cluster configuration, the others are pointed at it
                const std::string first = format_ipv4(hl[0].ipv4);
                if (&h == &hl[0]) inner.insert(inner.end(), {"-builtin-config-port", builtin_port});
                else if (config_server.empty()) inner.insert(inner.end(), {"-config-server", "http://" + first + ":" + builtin_port + "/config"});
            }
            if (!config_server.empty()) inner.insert(inner.end(), {"-config-server", config_server});
            if (!strategy.empty()) inner.insert(inner.end(), {"-strategy", strategy});
            if (!port.empty()) inner.insert(inner.end(), {"-port", port});
            if (!port_range.empty()) inner.insert(inner.end(), {"-port-range", port_range});
            if (!nic.empty()) inner.insert(inner.end(), {"-nic", nic});
            if (!logdir.empty()) inner.insert(inner.end(), {"-logdir", logdir});
            if (quiet) inner.push_back("-q");
            inner.insert(inner.end(), prog.begin(), prog.end());
        }
        Task t;
        t.name = ip;
        if (local) t.argv = inner;
        else t.argv = {"ssh", "-o", "StrictHostKeyChecking=no", "-o", "BatchMode=yes", (user.empty() ? "" : user + "@") + target, join_quoted(inner)};
        tasks.push_back(std::move(t));
    }
    if (dry) {
        for (auto& t : tasks) std::printf("%s\n", join_quoted(t.argv).c_str());
        return 0;
    }
    if (!logdir.empty()) mkdir(logdir.c_str(), 0755);
    const bool color = isatty(1);
    signal(SIGPIPE, SIG_IGN);
    for (auto& t : tasks) {
        if (!logdir.empty()) t.log = std::fopen((logdir + "/" + t.name + ".rrun.log").c_str(), "w");
        if (!spawn(t)) {
            std::fprintf(stderr, "kungfu-rrun: cannot start the task for %s: %s\n", t.name.c_str(), strerror(errno));
            return 1;
        }
    }
    // pump every child's output line by line; the first failure cancels the others (their whole process groups)
    int failed = 0;
    size_t live = tasks.size();
    bool cancelled = false;
    while (live > 0) {
        std::vector<pollfd> fds;
        std::vector<size_t> idx;
        for (size_t k = 0; k < tasks.size(); ++k)
            if (tasks[k].fd >= 0) fds.push_back({tasks[k].fd, POLLIN, 0}), idx.push_back(k);
        if (!fds.empty()) poll(fds.data(), fds.size(), 200);
        for (size_t f = 0; f < fds.size(); ++f) {
            Task& t = tasks[idx[f]];
            if (!(fds[f].revents & (POLLIN | POLLHUP | POLLERR))) continue;
            char buf[4096];
            const ssize_t n = read(t.fd, buf, sizeof(buf));
            if (n > 0) {
                t.partial.append(buf, (size_t)n);
                size_t pos;
                while ((pos = t.partial.find('\n')) != std::string::npos) {
                    emit(t, idx[f], t.partial.substr(0, pos), color);
                    t.partial.erase(0, pos + 1);
                }
                std::fflush(stdout);
            } else if (n == 0 || (n < 0 && errno != EINTR && errno != EAGAIN)) {
                if (!t.partial.empty()) emit(t, idx[f], t.partial, color);
                t.partial.clear();
                close(t.fd);
                t.fd = -1;
            }
        }
        for (auto& t : tasks) {
            if (t.done) continue;
            int st = 0;
            const pid_t r = waitpid(t.pid, &st, WNOHANG);
            if (r == t.pid && t.fd < 0) {
                t.done = true;
                --live;
                t.status = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
                if (t.status != 0) {
                    ++failed;
                    if (!cancelled) std::fprintf(stderr, "kungfu-rrun: %s exited with %d, cancelling the other hosts\n", t.name.c_str(), t.status);
                    if (!cancelled) {
                        cancelled = true;
                        for (auto& o : tasks)
                            if (!o.done && o.pid > 0) kill(-o.pid, SIGTERM); // exactly the process groups we created
                    }
                }
                if (t.log) std::fclose(t.log), t.log = nullptr;
            } else if (r == t.pid) {
                // exited but output not drained yet: remember the status, keep pumping
                t.status = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
                t.pid = -t.pid; // reaped
                // drain synchronously: the pipe's writers are gone or will be soon
                char buf[4096];
                ssize_t n;
                while ((n = read(t.fd, buf, sizeof(buf))) > 0) t.partial.append(buf, (size_t)n);
                size_t pos;
                size_t me = (size_t)(&t - &tasks[0]);
                while ((pos = t.partial.find('\n')) != std::string::npos) {
                    emit(t, me, t.partial.substr(0, pos), color);
                    t.partial.erase(0, pos + 1);
                }
                if (!t.partial.empty()) emit(t, me, t.partial, color);
                t.partial.clear();
                close(t.fd);
                t.fd = -1;
                t.done = true;
                --live;
                if (t.status != 0) {
                    ++failed;
                    if (!cancelled) {
                        std::fprintf(stderr, "kungfu-rrun: %s exited with %d, cancelling the other hosts\n", t.name.c_str(), t.status);
                        cancelled = true;
                        for (auto& o : tasks)
                            if (!o.done && o.pid > 0) kill(-o.pid, SIGTERM);
                    }
                }
                if (t.log) std::fclose(t.log), t.log = nullptr;
            }
        }
    }
    if (failed) std::fprintf(stderr, "kungfu-rrun: %d of %zu hosts failed\n", failed, tasks.size());
    return failed ? 1 : 0;
}
