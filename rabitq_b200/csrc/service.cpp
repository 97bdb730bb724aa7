// service.cpp -- twin of the reference's HTTP service contract (crates/service/src/main.rs:36-88, args.rs:5-21) over the C ABI.
//
// Same routes and JSON shapes:
//   GET  /  and  GET /health  -> "Ok"                                   (main.rs:32-34, 83-84)
//   GET  /metrics             -> Metrics::to_str()                      (main.rs:46-48, src/metrics.rs:30-41)
//   POST /query  {"query": [f32...], "top_k": u32, "probe": u32} -> {"ids": [u32...], "scores": [f32...]}   (main.rs:36-44, 55-66)
// Same flags: -d/--dir (saved index directory), -p/--port (9000); -b/--bucket, -k/--key, -c/--cache-dir are accepted and
// ignored: the reference reranks from an S3 + sqlite cache (crates/disk), here the base vectors are resident in HBM
// (SURVEY.md section 8f rank 3), so "cache miss" in /metrics is always 0.  Extra flags: --device N, --max-batch N,
// --batch-wait-us T, --threads N.
//
// What is new is MICRO-BATCHING: the reference runs one query per request on a tokio worker; a GPU wants batches.  Every
// connection thread parks its request in a queue; one batcher thread drains it -- everything that is waiting, up to
// --max-batch, after at most --batch-wait-us of lingering for company -- groups by (len, probe, top_k) and answers each
// group with ONE rabitq_query_batch call.  Results are exactly those of per-request rabitq_query calls (the batch entry is
// the same pipeline), ascending by distance (the reference returns heap order).
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <signal.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "rabitq_b200.h"

namespace {

std::atomic<bool> g_stop{false};
int g_listen_fd = -1;

void on_signal(int) {
    g_stop.store(true);
    if (g_listen_fd >= 0) ::shutdown(g_listen_fd, SHUT_RDWR);  // wakes accept()
}

#define INFO(...) do { std::fprintf(stderr, "[INFO  rabitq_service] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)

constexpr uint32_t kMaxTopK = 1024;          // results per request (validated before any buffer is sized from it)
constexpr size_t kMaxQueryLen = 65536;        // floats per query
constexpr size_t kMaxBodyBytes = 4u << 20;    // Content-Length cap: a query of 65536 floats in decimal is ~1 MB

struct Pending {  // one /query request waiting for its batch
    std::vector<float> query;
    uint32_t top_k = 0, probe = 0;
    std::vector<uint32_t> ids;
    std::vector<float> scores;
    int rc = 0;
    std::string err;
    bool done = false;
    std::mutex mu;
    std::condition_variable cv;
};

struct Batcher {
    rabitq_index* idx = nullptr;
    size_t max_batch = 1024;
    int wait_us = 200;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Pending*> queue;
    std::atomic<uint64_t> batches{0}, requests{0};

    void submit(Pending* p) {
        {
            std::lock_guard<std::mutex> lk(mu);
            queue.push_back(p);
        }
        cv.notify_one();
    }

    void run() {
        std::vector<Pending*> take;
        while (true) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !queue.empty() || g_stop.load(); });
                if (queue.empty() && g_stop.load()) return;
                if (queue.size() < max_batch && wait_us > 0) {  // linger briefly: concurrent clients fill the batch
                    cv.wait_for(lk, std::chrono::microseconds(wait_us), [&] { return queue.size() >= max_batch || g_stop.load(); });
                }
                take.clear();
                while (!queue.empty() && take.size() < max_batch) {
                    take.push_back(queue.front());
                    queue.pop_front();
                }
            }
            // one rabitq_query_batch call per (len, probe, top_k) group
            std::map<std::tuple<size_t, uint32_t, uint32_t>, std::vector<Pending*>> groups;
            for (Pending* p : take) groups[std::make_tuple(p->query.size(), p->probe, p->top_k)].push_back(p);
            for (auto& kv : groups) {
                const size_t len = std::get<0>(kv.first), probe = std::get<1>(kv.first), topk = std::get<2>(kv.first);
                auto& g = kv.second;
                const size_t nq = g.size();
                int rc = 0;
                std::string err;
                std::vector<float> dist;
                std::vector<uint32_t> ids, cnt;
                try {  // one bad group (e.g. an allocation failure) must not take the batcher thread -- and the process -- down
                    std::vector<float> q(nq * len);
                    dist.resize(nq * topk);
                    ids.resize(nq * topk);
                    cnt.resize(nq);
                    for (size_t i = 0; i < nq; i++) std::memcpy(q.data() + i * len, g[i]->query.data(), len * 4);
                    rc = rabitq_query_batch(idx, q.data(), nq, len, probe, topk, 0, dist.data(), ids.data(), cnt.data());
                    if (rc) err = rabitq_last_error();
                } catch (const std::exception& e) {
                    rc = -1;
                    err = std::string("batch failed: ") + e.what();
                }
                batches++;
                requests += nq;
                for (size_t i = 0; i < nq; i++) {
                    Pending* p = g[i];
                    std::lock_guard<std::mutex> lk(p->mu);
                    p->rc = rc;
                    p->err = err;
                    if (!rc) {
                        p->ids.assign(ids.begin() + i * topk, ids.begin() + i * topk + cnt[i]);
                        p->scores.assign(dist.begin() + i * topk, dist.begin() + i * topk + cnt[i]);
                    }
                    p->done = true;
                    // notified while the lock is held: `p` lives on the connection thread's stack, which may destroy it as
                    // soon as it can observe done == true
                    p->cv.notify_one();
                }
            }
        }
    }
};

// ---- the little JSON this contract needs (struct Request, main.rs:55-60) -----------------------------------------------
bool find_key(const std::string& body, const char* key, size_t* pos) {
    const std::string k = std::string("\"") + key + "\"";
    size_t p = body.find(k);
    if (p == std::string::npos) return false;
    p = body.find(':', p + k.size());
    if (p == std::string::npos) return false;
    *pos = p + 1;
    return true;
}

bool parse_request(const std::string& body, Pending* out, std::string* err) {
    size_t p;
    if (!find_key(body, "query", &p)) { *err = "missing field `query`"; return false; }
    p = body.find('[', p);
    if (p == std::string::npos) { *err = "`query` is not an array"; return false; }
    const char* s = body.c_str() + p + 1;
    while (true) {
        while (*s == ' ' || *s == '\n' || *s == '\r' || *s == '\t' || *s == ',') s++;
        if (*s == ']' || *s == 0) break;
        char* e = nullptr;
        float v = std::strtof(s, &e);
        if (e == s) { *err = "`query` holds a non-number"; return false; }
        out->query.push_back(v);
        s = e;
    }
    if (*s != ']') { *err = "unterminated `query` array"; return false; }
    auto get_u32 = [&](const char* key, uint32_t* v) {
        size_t q;
        if (!find_key(body, key, &q)) { *err = std::string("missing field `") + key + "`"; return false; }
        const char* b = body.c_str() + q;
        while (*b == ' ' || *b == '\t' || *b == '\n' || *b == '\r') b++;
        if (*b < '0' || *b > '9') { *err = std::string("`") + key + "` is not an unsigned integer"; return false; }  // u32 in the reference: '-1' is a deserialisation error there too
        unsigned long long x = 0;
        for (; *b >= '0' && *b <= '9'; b++) {
            x = x * 10 + (unsigned)(*b - '0');
            if (x > 0xffffffffull) { *err = std::string("`") + key + "` does not fit u32"; return false; }
        }
        *v = (uint32_t)x;
        return true;
    };
    if (!get_u32("top_k", &out->top_k) || !get_u32("probe", &out->probe)) return false;
    // the reference would panic (top_k == 0: empty heap peek; probe == 0: `length - 1` underflow) or allocate without bound; a
    // service must not let one request take the process down
    if (out->top_k == 0 || out->top_k > kMaxTopK) { *err = "`top_k` must be in 1.." + std::to_string(kMaxTopK); return false; }
    if (out->probe == 0) { *err = "`probe` must be >= 1"; return false; }
    if (out->query.empty() || out->query.size() > kMaxQueryLen) { *err = "`query` is empty or too long"; return false; }
    return true;
}

std::string to_json(const Pending& p) {  // struct Response, main.rs:62-66
    std::string s = "{\"ids\":[";
    char buf[40];
    for (size_t i = 0; i < p.ids.size(); i++) {
        std::snprintf(buf, sizeof buf, "%s%u", i ? "," : "", p.ids[i]);
        s += buf;
    }
    s += "],\"scores\":[";
    for (size_t i = 0; i < p.scores.size(); i++) {
        std::snprintf(buf, sizeof buf, "%s%.9g", i ? "," : "", (double)p.scores[i]);  // 9 digits round-trip an f32
        s += buf;
    }
    s += "]}";
    return s;
}

// ---- HTTP/1.1, just enough: request line, headers, Content-Length body, keep-alive ---------------------------------------
bool send_all(int fd, const char* p, size_t n) {
    while (n) {
        ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
        if (w <= 0) return false;
        p += w;
        n -= (size_t)w;
    }
    return true;
}

bool respond(int fd, int code, const char* reason, const char* ctype, const std::string& body, bool keep) {
    char head[256];
    int n = std::snprintf(head, sizeof head, "HTTP/1.1 %d %s\r\ncontent-type: %s\r\ncontent-length: %zu\r\nconnection: %s\r\n\r\n", code,
                          reason, ctype, body.size(), keep ? "keep-alive" : "close");
    return send_all(fd, head, (size_t)n) && send_all(fd, body.data(), body.size());
}

std::string lower(std::string s) {
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

std::string metrics_str(rabitq_index* idx) {  // Metrics::to_str, src/metrics.rs:30-41
    uint64_t m[4];
    rabitq_metrics(idx, m);
    char buf[256];
    std::snprintf(buf, sizeof buf, "query: %llu, rough: %llu, precise: %llu, ratio: %.2f, cache miss: %llu", (unsigned long long)m[0],
                  (unsigned long long)m[1], (unsigned long long)m[2], m[2] ? (double)m[1] / (double)m[2] : 0.0 / 0.0, (unsigned long long)m[3]);
    return buf;
}

void serve_connection(int fd, Batcher* b) {
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    std::string buf;
    char tmp[65536];
    while (!g_stop.load()) {
        size_t hdr_end;
        while ((hdr_end = buf.find("\r\n\r\n")) == std::string::npos) {
            ssize_t r = ::recv(fd, tmp, sizeof tmp, 0);
            if (r <= 0) { ::close(fd); return; }
            buf.append(tmp, (size_t)r);
            if (buf.size() > (64u << 20)) { ::close(fd); return; }
        }
        const std::string head = buf.substr(0, hdr_end);
        const size_t sp1 = head.find(' '), sp2 = head.find(' ', sp1 + 1);
        if (sp1 == std::string::npos || sp2 == std::string::npos) { respond(fd, 400, "Bad Request", "text/plain", "bad request line", false); break; }
        const std::string method = head.substr(0, sp1), path = head.substr(sp1 + 1, sp2 - sp1 - 1);
        const std::string lhead = lower(head);
        size_t clen = 0;
        {
            size_t p = lhead.find("\r\ncontent-length:");
            if (p != std::string::npos) clen = std::strtoul(lhead.c_str() + p + 17, nullptr, 10);
        }
        const bool keep = lhead.find("\r\nconnection: close") == std::string::npos;
        if (clen > kMaxBodyBytes) { respond(fd, 413, "Payload Too Large", "text/plain", "body too large", false); break; }
        while (buf.size() < hdr_end + 4 + clen) {
            ssize_t r = ::recv(fd, tmp, sizeof tmp, 0);
            if (r <= 0) { ::close(fd); return; }
            buf.append(tmp, (size_t)r);
        }
        const std::string body = buf.substr(hdr_end + 4, clen);
        buf.erase(0, hdr_end + 4 + clen);
        bool ok;
        if (method == "GET" && (path == "/" || path == "/health")) ok = respond(fd, 200, "OK", "text/plain; charset=utf-8", "Ok", keep);
        else if (method == "GET" && path == "/metrics") ok = respond(fd, 200, "OK", "text/plain; charset=utf-8", metrics_str(b->idx), keep);
        else if (method == "POST" && path == "/query") {
            Pending p;
            std::string err;
            if (!parse_request(body, &p, &err)) {
                ok = respond(fd, 422, "Unprocessable Entity", "text/plain; charset=utf-8", "Failed to deserialize the JSON body: " + err, keep);
            } else {
                b->submit(&p);
                {
                    std::unique_lock<std::mutex> lk(p.mu);
                    p.cv.wait(lk, [&] { return p.done; });
                }
                // the reference panics inside the handler on a bad query (assert at src/rabitq.rs:275) -> 500 for the client
                if (p.rc) ok = respond(fd, 500, "Internal Server Error", "text/plain; charset=utf-8", p.err, keep);
                else ok = respond(fd, 200, "OK", "application/json", to_json(p), keep);
            }
        } else if (path == "/" || path == "/health" || path == "/metrics" || path == "/query") {
            ok = respond(fd, 405, "Method Not Allowed", "text/plain", "", keep);
        } else {
            ok = respond(fd, 404, "Not Found", "text/plain", "", keep);
        }
        if (!ok || !keep) break;
    }
    ::close(fd);
}

[[noreturn]] void usage(const char* msg) {
    if (msg) std::fprintf(stderr, "%s\n", msg);
    std::fprintf(stderr,
                 "Usage: rabitq_service -d <dir> [-p <port>] [-b <bucket>] [-k <key>] [-c <cache-dir>] [--device N] [--max-batch N] "
                 "[--batch-wait-us T]\n\nRaBitQ\n\nOptions:\n  -d, --dir         the RaBitQ saved directory\n  -p, --port        service port (9000)\n"
                 "  -b, --bucket      S3 bucket (ignored: the base vectors are resident in HBM)\n  -k, --key         S3 key prefix (ignored)\n"
                 "  -c, --cache-dir   local cache directory (ignored)\n      --device      CUDA device ordinal (0)\n"
                 "      --max-batch   most requests answered by one rabitq_query_batch call (1024)\n"
                 "      --batch-wait-us  how long a request may wait for company (200)\n");
    std::exit(msg ? 2 : 0);
}

}  // namespace

int main(int argc, char** argv) {
    std::string dir;
    int port = 9000, device = 0;
    Batcher batcher;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&]() -> const char* {
            if (i + 1 >= argc) usage(("missing value for " + a).c_str());
            return argv[++i];
        };
        if (a == "-d" || a == "--dir") dir = val();
        else if (a == "-p" || a == "--port") port = std::atoi(val());
        else if (a == "-b" || a == "--bucket" || a == "-k" || a == "--key" || a == "-c" || a == "--cache-dir") (void)val();
        else if (a == "--device") device = std::atoi(val());
        else if (a == "--max-batch") batcher.max_batch = (size_t)std::max(1, std::atoi(val()));
        else if (a == "--batch-wait-us") batcher.wait_us = std::max(0, std::atoi(val()));
        else if (a == "--help") usage(nullptr);
        else usage(("unknown option " + a).c_str());
    }
    if (dir.empty()) usage("-d/--dir is required");
    if (rabitq_load_from_dir(dir.c_str(), device, &batcher.idx) != RABITQ_OK) {
        std::fprintf(stderr, "failed to load the index: %s\n", rabitq_last_error());  // the reference panics here (expect)
        return 1;
    }
    INFO("loaded %zu vectors, dim %u, %zu clusters onto CUDA device %d", rabitq_num_vectors(batcher.idx), rabitq_dim(batcher.idx),
         rabitq_num_clusters(batcher.idx), device);

    struct sigaction sa;
    std::memset(&sa, 0, sizeof sa);
    sa.sa_handler = on_signal;
    sigaction(SIGINT, &sa, nullptr);
    sigaction(SIGTERM, &sa, nullptr);

    g_listen_fd = ::socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(g_listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in addr;
    std::memset(&addr, 0, sizeof addr);
    addr.sin_family = AF_INET;
    addr.sin_addr.s_addr = htonl(INADDR_ANY);  // 0.0.0.0:{port}, main.rs:89
    addr.sin_port = htons((uint16_t)port);
    if (::bind(g_listen_fd, (sockaddr*)&addr, sizeof addr) != 0 || ::listen(g_listen_fd, 1024) != 0) {
        std::perror("bind/listen");
        return 1;
    }
    INFO("Server listening on 0.0.0.0:%d", port);
    std::thread bt([&] { batcher.run(); });
    std::vector<std::thread> conns;
    while (!g_stop.load()) {
        int fd = ::accept(g_listen_fd, nullptr, nullptr);
        if (fd < 0) {
            if (g_stop.load()) break;
            continue;
        }
        std::thread(serve_connection, fd, &batcher).detach();
    }
    INFO("Shutting down");
    g_stop.store(true);
    batcher.cv.notify_all();
    bt.join();
    ::close(g_listen_fd);
    INFO("answered %llu requests in %llu batches", (unsigned long long)batcher.requests.load(), (unsigned long long)batcher.batches.load());
    std::this_thread::sleep_for(std::chrono::milliseconds(50));  // let in-flight connection threads write their last bytes
    rabitq_free(batcher.idx);
    return 0;
}
