// voc_server -- native, protocol-compatible vocoder server on libvoc_b200.so.
//
// Speaks the wire protocol of /root/reference/dual_npu/vocoder_server.py:8-12,123-190 byte for byte:
//   client -> server : int32 LE n_tokens, then n_tokens * 16 int64 LE codes (row-major [n, 16])
//   server -> client : int32 LE n_samples, then n_samples int16 LE PCM
//   n_tokens <= 0 or > 10000, a short body, or any synthesis error: the connection is closed
//   without a reply (:149-151,162-164,180-183).  Socket default /tmp/qwen3_voc.sock, mode 0666 (:131,196).
// The reference serves one connection at a time (listen(1), :129); its streaming client, however, opens
// one connection per 64 accumulated tokens from separate threads (dual_npu/tts_client.py:188-197).
//
// Two threads.  The I/O thread owns every socket: a poll() loop that accepts, reads request bytes and writes
// reply bytes, all non-blocking, so a slow or stalled peer never holds up anybody else.  The GPU thread owns
// the vocoder handle: it takes every request that is complete (up to --max-batch), synthesises them with ONE
// voc_synthesize_batch_pcm16 call (their windows share batched launches and one stitch; each reply is
// bit-identical to serving that request alone) and hands the replies back.  While the GPU works on one batch
// the I/O thread keeps reading the next requests and draining the previous replies.
// Bounds: at most --max-conns connections (further ones are accepted and closed at once), a request must
// arrive within --recv-timeout-ms and its reply must be taken within --send-timeout-ms, else the connection
// is dropped; accept() failing for lack of descriptors backs off instead of spinning.
// Like the reference's own native servers (dual_npu/code_predictor_cpp/code_predictor_server.cpp:422-558):
// plain C++17, POSIX sockets, no framework.
//
//   voc_server --model vocoder.b200voc [--socket /tmp/qwen3_voc.sock] [--device 0] [--wave 32]
//              [--max-batch 64] [--window-us 500] [--max-conns 256] [--recv-timeout-ms 5000]
//              [--send-timeout-ms 10000]
#include "../../include/voc_b200.h"

#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <list>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

namespace {

constexpr int kMaxRequestTokens = 10000;        // vocoder_server.py:149
constexpr int kCodebooks = 16;

volatile sig_atomic_t g_running = 1;
void on_signal(int) { g_running = 0; }

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

enum State { READING, QUEUED, WRITING, DEAD };

struct Conn {
    int fd = -1;
    uint64_t id = 0;
    State st = READING;
    unsigned char hdr[4];
    size_t got = 0;                             // bytes of header + body received
    int n_tokens = 0;
    std::vector<long long> codes;               // the body, reserved once the header is known
    std::vector<unsigned char> out;             // the reply: int32 n_samples + PCM
    size_t sent = 0;
    double deadline = 0;                        // receive or send deadline
};

struct Job { uint64_t id; int n_tokens; std::vector<long long> codes; };
struct Done { uint64_t id; bool ok; std::vector<unsigned char> reply; };

struct Shared {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> pending;
    std::deque<Done> done;
    std::atomic<int> reading{0};                // connections whose request is still arriving
    bool stop = false;
    int wake_fd = -1;                           // write end of the I/O thread's self-pipe
};

// non-blocking read of whatever is available
void pump(Conn& c) {
    while (c.st == READING) {
        unsigned char* dst; size_t want;
        if (c.got < 4) { dst = c.hdr + c.got; want = 4 - c.got; }
        else {
            const size_t body = (size_t)c.n_tokens * kCodebooks * 8;
            dst = (unsigned char*)c.codes.data() + (c.got - 4); want = body - (c.got - 4);
        }
        const ssize_t r = recv(c.fd, dst, want, MSG_DONTWAIT);
        if (r == 0) { c.st = DEAD; return; }                               // peer closed early: no reply
        if (r < 0) {
            if (errno == EAGAIN || errno == EWOULDBLOCK) return;
            if (errno == EINTR) continue;
            c.st = DEAD; return;
        }
        c.got += (size_t)r;
        if (c.got == 4) {
            int32_t n;
            memcpy(&n, c.hdr, 4);                                          // little-endian host
            if (n <= 0 || n > kMaxRequestTokens) { c.st = DEAD; return; }  // :149-151
            c.n_tokens = n;
            c.codes.resize((size_t)n * kCodebooks);                        // one allocation per request
        } else if (c.got > 4 && c.got == 4 + (size_t)c.n_tokens * kCodebooks * 8) {
            c.st = QUEUED;
        }
    }
}

// non-blocking write of as much of the reply as the socket takes
void drain(Conn& c) {
    while (c.st == WRITING && c.sent < c.out.size()) {
        const ssize_t w = send(c.fd, c.out.data() + c.sent, c.out.size() - c.sent, MSG_NOSIGNAL | MSG_DONTWAIT);
        if (w < 0) {
            if (errno == EAGAIN || errno == EWOULDBLOCK) return;
            if (errno == EINTR) continue;
            c.st = DEAD; return;
        }
        c.sent += (size_t)w;
    }
    if (c.st == WRITING && c.sent == c.out.size()) c.st = DEAD;           // reply complete: close
}

void gpu_thread(void* voc, Shared* S, int max_batch, int window_us) {
    std::vector<long long> codes;
    std::vector<int> lens;
    std::vector<short> pcm;
    std::vector<long long> offs;
    for (;;) {
        std::vector<Job> batch;
        {
            std::unique_lock<std::mutex> lk(S->mu);
            S->cv.wait(lk, [&] { return S->stop || !S->pending.empty(); });
            if (S->stop) return;
            // a short coalescing window: requests still arriving on other connections join the batch
            if (window_us > 0) {
                const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(window_us);
                while ((int)S->pending.size() < max_batch && S->reading.load() > 0 && !S->stop)
                    if (S->cv.wait_until(lk, t_end) == std::cv_status::timeout) break;
                if (S->stop) return;
            }
            while (!S->pending.empty() && (int)batch.size() < max_batch) {
                batch.push_back(std::move(S->pending.front()));
                S->pending.pop_front();
            }
        }
        const double t0 = now_s();
        codes.clear(); lens.clear();
        long long cap = 0;
        for (auto& j : batch) {
            codes.insert(codes.end(), j.codes.begin(), j.codes.end());
            lens.push_back(j.n_tokens);
            cap += voc_out_samples(voc, j.n_tokens);
        }
        pcm.resize((size_t)cap);
        offs.assign(batch.size() + 1, 0);
        std::vector<Done> out;
        auto make_reply = [](const short* p, long long n) {
            std::vector<unsigned char> r(4 + (size_t)n * 2);
            const int32_t n32 = (int32_t)n;
            memcpy(r.data(), &n32, 4);
            memcpy(r.data() + 4, p, (size_t)n * 2);
            return r;
        };
        const int err = voc_synthesize_batch_pcm16(voc, codes.data(), lens.data(), (int)batch.size(), pcm.data(), cap, offs.data());
        if (err == VOC_OK) {
            long long tok = 0;
            for (size_t k = 0; k < batch.size(); ++k) {
                out.push_back({batch[k].id, true, make_reply(pcm.data() + offs[k], offs[k + 1] - offs[k])});
                tok += lens[k];
            }
            printf("  Vocoder: %lld tokens -> %lld samples (%.2fs) [%zu request%s]\n", tok, offs[batch.size()],
                   now_s() - t0, batch.size(), batch.size() == 1 ? "" : "s");
        } else {
            // one bad request (e.g. a code outside [0, 2048)) must not take the others down: serve them one by
            // one; the failing ones are closed without a reply, as the reference does on any exception
            for (auto& j : batch) {
                long long n_out = 0;
                const long long cap1 = voc_out_samples(voc, j.n_tokens);
                pcm.resize((size_t)cap1);
                const int e1 = voc_synthesize_pcm16(voc, j.codes.data(), j.n_tokens, pcm.data(), cap1, &n_out);
                if (e1 == VOC_OK) {
                    out.push_back({j.id, true, make_reply(pcm.data(), n_out)});
                    printf("  Vocoder: %d tokens -> %lld samples (%.2fs)\n", j.n_tokens, n_out, now_s() - t0);
                } else {
                    out.push_back({j.id, false, {}});
                    printf("  Vocoder Error: %s\n", voc_last_error(voc));
                }
            }
        }
        fflush(stdout);
        {
            std::lock_guard<std::mutex> lk(S->mu);
            for (auto& d : out) S->done.push_back(std::move(d));
        }
        const char b = 1;
        if (write(S->wake_fd, &b, 1) < 0) { /* the pipe is full: the I/O thread is awake anyway */ }
    }
}

}  // namespace

int main(int argc, char** argv) {
    std::string model, sock_path = "/tmp/qwen3_voc.sock";
    int device = 0, wave = 32, max_batch = 64, window_us = 500, max_conns = 256, recv_ms = 5000, send_ms = 10000;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--model") model = next();
        else if (a == "--socket") sock_path = next();
        else if (a == "--device") device = atoi(next());
        else if (a == "--wave") wave = atoi(next());
        else if (a == "--max-batch") max_batch = atoi(next());
        else if (a == "--window-us") window_us = atoi(next());
        else if (a == "--max-conns") max_conns = atoi(next());
        else if (a == "--recv-timeout-ms") recv_ms = atoi(next());
        else if (a == "--send-timeout-ms") send_ms = atoi(next());
        else {
            fprintf(stderr, "usage: voc_server --model M.b200voc [--socket P] [--device D] [--wave W] [--max-batch B] "
                            "[--window-us U] [--max-conns C] [--recv-timeout-ms R] [--send-timeout-ms S]\n");
            return 2;
        }
    }
    if (model.empty()) { fprintf(stderr, "voc_server: --model is required\n"); return 2; }
    if (max_batch < 1) max_batch = 1;
    if (max_conns < 1) max_conns = 1;

    signal(SIGPIPE, SIG_IGN);
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_handler = on_signal;
    sigaction(SIGINT, &sa, nullptr);
    sigaction(SIGTERM, &sa, nullptr);

    void* voc = voc_create_from_file(model.c_str(), device, wave);
    if (!voc) { fprintf(stderr, "voc_server: %s\n", voc_last_error(nullptr)); return 1; }
    printf("Vocoder: B200 CUDA (device %d), max_tokens=%d\n", device, voc_max_tokens(voc));

    unlink(sock_path.c_str());
    const int lfd = socket(AF_UNIX, SOCK_STREAM, 0);
    sockaddr_un addr;
    memset(&addr, 0, sizeof addr);
    addr.sun_family = AF_UNIX;
    strncpy(addr.sun_path, sock_path.c_str(), sizeof(addr.sun_path) - 1);
    int wake[2] = {-1, -1};
    if (lfd < 0 || bind(lfd, (sockaddr*)&addr, sizeof addr) < 0 || listen(lfd, 128) < 0 || pipe(wake) < 0) {
        perror("voc_server: socket/bind/listen");
        voc_destroy(voc);
        return 1;
    }
    chmod(sock_path.c_str(), 0666);
    fcntl(lfd, F_SETFL, fcntl(lfd, F_GETFL, 0) | O_NONBLOCK);
    fcntl(wake[0], F_SETFL, fcntl(wake[0], F_GETFL, 0) | O_NONBLOCK);
    fcntl(wake[1], F_SETFL, fcntl(wake[1], F_GETFL, 0) | O_NONBLOCK);
    printf("\nVocoder Server listening on %s\n", sock_path.c_str());
    fflush(stdout);

    Shared S;
    S.wake_fd = wake[1];
    std::thread gpu(gpu_thread, voc, &S, max_batch, window_us);

    std::list<Conn> conns;
    uint64_t next_id = 1;
    double accept_backoff_until = 0;
    std::vector<pollfd> pfds;
    std::vector<Conn*> who;

    while (g_running) {
        // ---- wait for traffic (1 s tick so that a signal is noticed, like the reference's settimeout(1.0))
        const double t = now_s();
        pfds.clear(); who.clear();
        pfds.push_back({wake[0], POLLIN, 0}); who.push_back(nullptr);
        const bool poll_listen = t >= accept_backoff_until;
        if (poll_listen) { pfds.push_back({lfd, POLLIN, 0}); who.push_back(nullptr); }
        double next_deadline = t + 1.0;
        for (auto& c : conns) {
            if (c.st == READING) { pfds.push_back({c.fd, POLLIN, 0}); who.push_back(&c); }
            else if (c.st == WRITING) { pfds.push_back({c.fd, POLLOUT, 0}); who.push_back(&c); }
            if ((c.st == READING || c.st == WRITING) && c.deadline < next_deadline) next_deadline = c.deadline;
        }
        if (!poll_listen && accept_backoff_until < next_deadline) next_deadline = accept_backoff_until;
        int timeout_ms = (int)((next_deadline - t) * 1000.0) + 1;
        if (timeout_ms < 1) timeout_ms = 1;
        const int rc = poll(pfds.data(), (nfds_t)pfds.size(), timeout_ms);
        if (rc < 0 && errno != EINTR) { perror("voc_server: poll"); break; }
        const double now = now_s();

        // ---- replies handed back by the GPU thread
        { char junk[64]; while (read(wake[0], junk, sizeof junk) > 0) {} }
        {
            std::deque<Done> done;
            { std::lock_guard<std::mutex> lk(S.mu); done.swap(S.done); }
            for (auto& d : done) {
                for (auto& c : conns) {
                    if (c.id != d.id) continue;
                    if (c.st == QUEUED) {
                        if (d.ok) { c.out = std::move(d.reply); c.sent = 0; c.st = WRITING; c.deadline = now + send_ms * 1e-3; drain(c); }
                        else c.st = DEAD;                                 // synthesis error: close without a reply (:180-183)
                    }
                    break;
                }
            }
        }
        // ---- accept
        if (poll_listen) {
            for (;;) {
                const int fd = accept(lfd, nullptr, nullptr);
                if (fd < 0) {
                    if (errno == EMFILE || errno == ENFILE || errno == ENOBUFS || errno == ENOMEM) {
                        // out of descriptors: lfd stays readable, so polling it again at once would spin
                        fprintf(stderr, "voc_server: accept: %s; backing off\n", strerror(errno));
                        accept_backoff_until = now + 0.05;
                    } else if (errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR && errno != ECONNABORTED) {
                        perror("voc_server: accept");
                    }
                    break;
                }
                if ((int)conns.size() >= max_conns) { close(fd); continue; }      // over the cap: refused
                fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK);
                conns.emplace_back();
                Conn& c = conns.back();
                c.fd = fd; c.id = next_id++; c.deadline = now + recv_ms * 1e-3;
                S.reading.fetch_add(1);
                pump(c);
                if (c.st != READING) S.reading.fetch_sub(1);
            }
        }
        // ---- read / write whatever poll reported
        for (size_t i = 0; i < pfds.size(); ++i) {
            Conn* c = who[i];
            if (!c || !pfds[i].revents) continue;
            if (c->st == READING) { pump(*c); if (c->st != READING) S.reading.fetch_sub(1); }
            else if (c->st == WRITING) drain(*c);
        }
        // ---- complete requests go to the GPU thread; expired and finished connections are closed
        bool queued = false;
        for (auto it = conns.begin(); it != conns.end();) {
            Conn& c = *it;
            if (c.st == QUEUED && !c.codes.empty()) {
                std::lock_guard<std::mutex> lk(S.mu);
                S.pending.push_back({c.id, c.n_tokens, std::move(c.codes)});
                c.codes.clear();
                queued = true;
            }
            if (c.st == READING && now > c.deadline) { c.st = DEAD; S.reading.fetch_sub(1); }   // stalled sender
            if (c.st == WRITING && now > c.deadline) c.st = DEAD;                             // stalled receiver
            if (c.st == DEAD) { close(c.fd); it = conns.erase(it); } else ++it;
        }
        if (queued || S.reading.load() == 0) S.cv.notify_all();
    }

    {
        std::lock_guard<std::mutex> lk(S.mu);
        S.stop = true;
    }
    S.cv.notify_all();
    gpu.join();
    for (auto& c : conns) if (c.fd >= 0) close(c.fd);
    close(lfd); close(wake[0]); close(wake[1]);
    unlink(sock_path.c_str());
    voc_destroy(voc);
    printf("Vocoder Server stopped.\n");
    return 0;
}
