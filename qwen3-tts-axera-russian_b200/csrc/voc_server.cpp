// voc_server -- native, protocol-compatible vocoder server on libvoc_b200.so.
//
// Speaks the wire protocol of /root/reference/dual_npu/vocoder_server.py:8-12,123-190 byte for byte:
//   client -> server : int32 LE n_tokens, then n_tokens * 16 int64 LE codes (row-major [n, 16])
//   server -> client : int32 LE n_samples, then n_samples int16 LE PCM
//   n_tokens <= 0 or > 10000, a short body, or any synthesis error: the connection is closed
//   without a reply (:149-151,162-164,180-183).  Socket default /tmp/qwen3_voc.sock, mode 0666 (:131,196).
// The reference serves one connection at a time (listen(1), :129); its streaming client, however, opens
// one connection per 64 accumulated tokens from separate threads (dual_npu/tts_client.py:188-197).  This
// server keeps every connection open concurrently (poll), and all requests that are complete when the GPU
// becomes free are synthesised by ONE voc_synthesize_batch_pcm16 call: their windows share batched
// launches and one stitch, and each reply is bit-identical to serving that request alone.
// Like the reference's own native servers (dual_npu/code_predictor_cpp/code_predictor_server.cpp:422-558):
// plain C++17, POSIX sockets, no framework.
//
//   voc_server --model vocoder.b200voc [--socket /tmp/qwen3_voc.sock] [--device 0] [--wave 32]
//              [--max-batch 64] [--window-us 500]
#include "../../include/voc_b200.h"

#include <cerrno>
#include <chrono>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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

struct Conn {
    int fd = -1;
    std::vector<unsigned char> buf;             // bytes received so far
    size_t want = 4;                            // header first, then header + body
    int n_tokens = 0;
    bool ready = false;                         // a complete request is waiting for the GPU
    bool dead = false;
};

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// non-blocking read of whatever is available; marks the connection ready / dead
void pump(Conn& c) {
    while (!c.ready && !c.dead) {
        if (c.buf.size() < c.want) {
            const size_t old = c.buf.size();
            c.buf.resize(c.want);
            const ssize_t r = recv(c.fd, c.buf.data() + old, c.want - old, MSG_DONTWAIT);
            if (r > 0) { c.buf.resize(old + (size_t)r); continue; }
            c.buf.resize(old);
            if (r == 0) { c.dead = true; return; }                        // peer closed early: no reply
            if (errno == EAGAIN || errno == EWOULDBLOCK) return;
            if (errno == EINTR) continue;
            c.dead = true; return;
        }
        if (c.want == 4) {
            int32_t n;
            memcpy(&n, c.buf.data(), 4);                                   // little-endian host
            if (n <= 0 || n > kMaxRequestTokens) { c.dead = true; return; }  // :149-151
            c.n_tokens = n;
            c.want = 4 + (size_t)n * kCodebooks * 8;
        } else {
            c.ready = true;
        }
    }
}

bool send_all(int fd, const void* data, size_t n) {
    const unsigned char* p = (const unsigned char*)data;
    while (n > 0) {
        const ssize_t w = send(fd, p, n, MSG_NOSIGNAL);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        p += w; n -= (size_t)w;
    }
    return true;
}

void reply_and_close(Conn& c, const short* pcm, long long n_samples) {
    // the reply is sent with the descriptor back in blocking mode, like the reference's sendall
    const int fl = fcntl(c.fd, F_GETFL, 0);
    fcntl(c.fd, F_SETFL, fl & ~O_NONBLOCK);
    const int32_t n = (int32_t)n_samples;
    if (send_all(c.fd, &n, 4)) send_all(c.fd, pcm, (size_t)n_samples * 2);
    close(c.fd);
    c.fd = -1;
}

}  // namespace

int main(int argc, char** argv) {
    std::string model, sock_path = "/tmp/qwen3_voc.sock";
    int device = 0, wave = 32, max_batch = 64, window_us = 500;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--model") model = next();
        else if (a == "--socket") sock_path = next();
        else if (a == "--device") device = atoi(next());
        else if (a == "--wave") wave = atoi(next());
        else if (a == "--max-batch") max_batch = atoi(next());
        else if (a == "--window-us") window_us = atoi(next());
        else { fprintf(stderr, "usage: voc_server --model M.b200voc [--socket P] [--device D] [--wave W] [--max-batch B] [--window-us U]\n"); return 2; }
    }
    if (model.empty()) { fprintf(stderr, "voc_server: --model is required\n"); return 2; }
    if (max_batch < 1) max_batch = 1;

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
    if (lfd < 0 || bind(lfd, (sockaddr*)&addr, sizeof addr) < 0 || listen(lfd, 128) < 0) {
        perror("voc_server: socket/bind/listen");
        voc_destroy(voc);
        return 1;
    }
    chmod(sock_path.c_str(), 0666);
    fcntl(lfd, F_SETFL, fcntl(lfd, F_GETFL, 0) | O_NONBLOCK);
    printf("\nVocoder Server listening on %s\n", sock_path.c_str());
    fflush(stdout);

    std::vector<Conn> conns;
    std::vector<long long> codes;
    std::vector<int> lens;
    std::vector<short> pcm;
    std::vector<long long> offs;

    while (g_running) {
        // ---- wait for traffic (1 s tick so that a signal is noticed, like the reference's settimeout(1.0))
        std::vector<pollfd> pfds;
        pfds.push_back({lfd, POLLIN, 0});
        for (auto& c : conns) pfds.push_back({c.fd, POLLIN, 0});
        int n_ready = 0;
        for (auto& c : conns) n_ready += c.ready;
        const int rc = poll(pfds.data(), (nfds_t)pfds.size(), n_ready ? 0 : 1000);
        if (rc < 0 && errno != EINTR) { perror("voc_server: poll"); break; }
        // ---- accept, read
        for (;;) {
            const int fd = accept(lfd, nullptr, nullptr);
            if (fd < 0) break;
            fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK);
            Conn c; c.fd = fd;
            conns.push_back(std::move(c));
        }
        for (auto& c : conns) pump(c);
        // drop connections that ended or sent a bad header, without a reply
        for (size_t i = 0; i < conns.size();) {
            if (conns[i].dead) { close(conns[i].fd); conns.erase(conns.begin() + (long)i); } else ++i;
        }
        n_ready = 0;
        for (auto& c : conns) n_ready += c.ready;
        if (!n_ready) continue;
        // ---- a short coalescing window: requests still in flight on other connections join the batch
        if (window_us > 0 && n_ready < (int)conns.size() && n_ready < max_batch) {
            const double t_end = now_s() + window_us * 1e-6;
            while (now_s() < t_end) {
                for (;;) {
                    const int fd = accept(lfd, nullptr, nullptr);
                    if (fd < 0) break;
                    fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK);
                    Conn c; c.fd = fd;
                    conns.push_back(std::move(c));
                }
                int pending = 0;
                for (auto& c : conns) { pump(c); pending += (!c.ready && !c.dead); }
                if (!pending) break;
                usleep(20);
            }
        }
        // ---- one batched synthesis for everything that is complete
        std::vector<size_t> batch;
        for (size_t i = 0; i < conns.size() && (int)batch.size() < max_batch; ++i)
            if (conns[i].ready && !conns[i].dead) batch.push_back(i);
        if (batch.empty()) continue;
        const double t0 = now_s();
        codes.clear(); lens.clear();
        long long cap = 0;
        for (size_t i : batch) {
            const Conn& c = conns[i];
            const size_t n_words = (size_t)c.n_tokens * kCodebooks;
            const size_t at = codes.size();
            codes.resize(at + n_words);
            memcpy(codes.data() + at, c.buf.data() + 4, n_words * 8);
            lens.push_back(c.n_tokens);
            cap += voc_out_samples(voc, c.n_tokens);
        }
        pcm.resize((size_t)cap);
        offs.assign(batch.size() + 1, 0);
        int err = voc_synthesize_batch_pcm16(voc, codes.data(), lens.data(), (int)batch.size(), pcm.data(), cap, offs.data());
        if (err == VOC_OK) {
            long long tok = 0;
            for (size_t k = 0; k < batch.size(); ++k) {
                reply_and_close(conns[batch[k]], pcm.data() + offs[k], offs[k + 1] - offs[k]);
                tok += lens[k];
            }
            printf("  Vocoder: %lld tokens -> %lld samples (%.2fs) [%zu request%s]\n", tok, offs[batch.size()],
                   now_s() - t0, batch.size(), batch.size() == 1 ? "" : "s");
        } else {
            // one bad request (e.g. a code outside [0, 2048)) must not take the others down: serve them one by
            // one; the failing ones are closed without a reply, as the reference does on any exception
            size_t at = 0;
            for (size_t k = 0; k < batch.size(); ++k) {
                Conn& c = conns[batch[k]];
                long long n_out = 0;
                const long long cap1 = voc_out_samples(voc, c.n_tokens);
                pcm.resize((size_t)cap1);
                const int e1 = voc_synthesize_pcm16(voc, codes.data() + at, c.n_tokens, pcm.data(), cap1, &n_out);
                at += (size_t)c.n_tokens * kCodebooks;
                if (e1 == VOC_OK) {
                    reply_and_close(c, pcm.data(), n_out);
                    printf("  Vocoder: %d tokens -> %lld samples (%.2fs)\n", c.n_tokens, n_out, now_s() - t0);
                } else {
                    printf("  Vocoder Error: %s\n", voc_last_error(voc));
                    close(c.fd); c.fd = -1;
                }
            }
        }
        fflush(stdout);
        for (size_t i = 0; i < conns.size();) {
            if (conns[i].fd < 0) conns.erase(conns.begin() + (long)i); else ++i;
        }
    }

    for (auto& c : conns) if (c.fd >= 0) close(c.fd);
    close(lfd);
    unlink(sock_path.c_str());
    voc_destroy(voc);
    printf("Vocoder Server stopped.\n");
    return 0;
}
