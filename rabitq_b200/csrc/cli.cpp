// cli.cpp -- twin of the reference's benchmark CLI (crates/cli/src/main.rs:11-83) over the C ABI.
//
// Same flags (-b/--base -c/--centroids -q/--query -t/--truth -p/--probe (100) -k/--topk (10) -s/--saved
// -h/--heuristic-rank), same log lines ("QPS: {}, recall: {}", "Metrics [...]", main.rs:77-82).  The index is
// loaded from the -s directory (RaBitQ::load_from_dir) or, when that does not exist, trained on the device from -b/-c
// (RaBitQ::from_path) and saved there (dump_to_dir), main.rs:51-61.  Extra flags: --device N, --seed S (the random
// rotation), --single (one rabitq_query call per query like the reference's loop, main.rs:69-75; default is one
// rabitq_query_batch call).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "rabitq_b200.h"

namespace {

template <typename T>
bool read_vecs(const std::string& path, std::vector<std::vector<T>>& out) {  // src/utils.rs:280-303
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    uint32_t dim;
    while (std::fread(&dim, 4, 1, f) == 1) {
        std::vector<T> v(dim);
        if (dim && std::fread(v.data(), sizeof(T), dim, f) != dim) { std::fclose(f); return false; }
        out.emplace_back(std::move(v));
    }
    std::fclose(f);
    return true;
}

// src/utils.rs:367-379
float calculate_recall(const std::vector<int32_t>& truth, const uint32_t* res, size_t topk) {
    int count = 0;
    for (size_t i = 0; i < topk; i++)
        for (size_t t = 0; t < topk && t < truth.size(); t++)
            if ((int32_t)res[i] == truth[t]) { count++; break; }
    return float(count) / float(topk);
}

bool is_dir(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

int level() {  // env_logger: RABITQ_LOG, default "debug" (main.rs:41)
    const char* e = std::getenv("RABITQ_LOG");
    std::string s = e ? e : "debug";
    if (s == "off" || s == "error" || s == "warn") return 0;
    if (s == "info") return 1;
    return 2;
}
#define DEBUG(...) do { if (level() >= 2) { std::fprintf(stderr, "[DEBUG rabitq_cli] " __VA_ARGS__); std::fprintf(stderr, "\n"); } } while (0)
#define INFO(...) do { if (level() >= 1) { std::fprintf(stderr, "[INFO  rabitq_cli] " __VA_ARGS__); std::fprintf(stderr, "\n"); } } while (0)

[[noreturn]] void die(const char* what) {
    std::fprintf(stderr, "%s: %s\n", what, rabitq_last_error());
    std::exit(101);  // a Rust panic exits with 101
}

}  // namespace

int main(int argc, char** argv) {
    std::string base, centroids, query, truth, saved;
    size_t probe = 100, topk = 10;
    bool heuristic = false, single = false;
    int device = 0;
    uint64_t seed = 42;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* name) -> std::string {
            if (i + 1 >= argc) { std::fprintf(stderr, "No value provided for option '%s'.\n", name); std::exit(1); }
            return argv[++i];
        };
        if (a == "-b" || a == "--base") base = val("--base");
        else if (a == "-c" || a == "--centroids") centroids = val("--centroids");
        else if (a == "-q" || a == "--query") query = val("--query");
        else if (a == "-t" || a == "--truth") truth = val("--truth");
        else if (a == "-p" || a == "--probe") probe = std::strtoull(val("--probe").c_str(), nullptr, 10);
        else if (a == "-k" || a == "--topk") topk = std::strtoull(val("--topk").c_str(), nullptr, 10);
        else if (a == "-s" || a == "--saved") saved = val("--saved");
        else if (a == "-h" || a == "--heuristic-rank") heuristic = true;
        else if (a == "--device") device = std::atoi(val("--device").c_str());
        else if (a == "--single") single = true;
        else if (a == "--seed") seed = std::strtoull(val("--seed").c_str(), nullptr, 10);
        else if (a == "--help") {
            std::printf("Usage: rabitq_cli -b <base> -c <centroids> -q <query> -t <truth> [-p <probe>] [-k <topk>] -s <saved> [-h]\n\n"
                        "RaBitQ CLI args (B200 twin of crates/cli)\n");
            return 0;
        } else { std::fprintf(stderr, "Unrecognized argument: %s\n", a.c_str()); return 1; }
    }
    if (query.empty() || truth.empty() || saved.empty()) {
        std::fprintf(stderr, "Required options not provided:\n    --query\n    --truth\n    --saved\n");
        return 1;
    }
    DEBUG("Args { base: \"%s\", centroids: \"%s\", query: \"%s\", truth: \"%s\", probe: %zu, topk: %zu, saved: \"%s\", heuristic_rank: %s }",
          base.c_str(), centroids.c_str(), query.c_str(), truth.c_str(), probe, topk, saved.c_str(), heuristic ? "true" : "false");
    rabitq_index* ix = nullptr;
    if (is_dir(saved)) {
        DEBUG("loading from \"%s\"...", saved.c_str());
        if (rabitq_load_from_dir(saved.c_str(), device, &ix)) die("load_from_dir");
    } else {  // crates/cli/src/main.rs:56-61
        if (base.empty() || centroids.empty()) {
            std::fprintf(stderr, "Required options not provided:\n    --base\n    --centroids\n");
            return 1;
        }
        DEBUG("training...");
        if (rabitq_from_path(base.c_str(), centroids.c_str(), seed, device, &ix)) die("from_path");
        DEBUG("saving to local file: \"%s\"", saved.c_str());
        if (rabitq_dump_to_dir(ix, saved.c_str())) die("dump_to_dir");
    }
    std::vector<std::vector<float>> queries;
    std::vector<std::vector<int32_t>> truths;
    if (!read_vecs(query, queries)) { std::fprintf(stderr, "read query error\n"); return 101; }
    if (!read_vecs(truth, truths)) { std::fprintf(stderr, "read truth error\n"); return 101; }
    DEBUG("querying...");
    const size_t nq = queries.size();
    if (nq == 0 || truths.size() < nq) { std::fprintf(stderr, "empty queries or too few truth rows\n"); return 101; }
    const size_t len = queries[0].size();
    std::vector<float> flat(nq * len);
    for (size_t i = 0; i < nq; i++) {
        // the reference asserts the length of every query inside RaBitQ::query (src/rabitq.rs:275); a ragged file must not be read out of bounds
        if (queries[i].size() != len) { std::fprintf(stderr, "query %zu has %zu dimensions, expected %zu\n", i, queries[i].size(), len); return 101; }
        std::memcpy(&flat[i * len], queries[i].data(), len * 4);
    }
    std::vector<float> dist(nq * topk);
    std::vector<uint32_t> ids(nq * topk), cnt(nq);
    double total_time = 0.0;
    if (single) {
        for (size_t i = 0; i < nq; i++) {
            auto t0 = std::chrono::steady_clock::now();
            if (rabitq_query(ix, &flat[i * len], len, probe, topk, heuristic, &dist[i * topk], &ids[i * topk], &cnt[i])) die("query");
            total_time += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
    } else {
        auto t0 = std::chrono::steady_clock::now();
        if (rabitq_query_batch(ix, flat.data(), nq, len, probe, topk, heuristic, dist.data(), ids.data(), cnt.data())) die("query");
        total_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    float recall = 0.f;
    for (size_t i = 0; i < nq; i++) {
        if (cnt[i] != topk) { std::fprintf(stderr, "assertion `left == right` failed: res.len() == topk (query %zu)\n", i); return 101; }
        recall += calculate_recall(truths[i], &ids[i * topk], topk);
    }
    INFO("QPS: %g, recall: %g", double(nq) / total_time, recall / float(nq));
    uint64_t m[4];
    rabitq_metrics(ix, m);
    INFO("Metrics [query: %llu, rough: %llu, precise: %llu, ratio: %.2f, cache miss: %llu]", (unsigned long long)m[0],
         (unsigned long long)m[1], (unsigned long long)m[2], double(m[1]) / double(m[2]), (unsigned long long)m[3]);
    rabitq_free(ix);
    return 0;
}
