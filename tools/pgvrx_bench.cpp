// pgvrx_bench.cpp -- standalone host-side bench binary over the C ABI (include/pgvrx_b200.h).
//
// SURVEY section 8(b) "what calls it": the north star's host side is a Rust bench binary over a thin FFI crate; there
// is no Rust toolchain in this image, so this is the same program in C++17.  It links against libpgvrx_b200.so and
// nothing else -- no Python, no torch, no CUDA headers: everything it does goes through the entry points a pgrx-sys /
// cgo binding would see.
//
//   g++ -O2 -std=c++17 -pthread -Iinclude tools/pgvrx_bench.cpp -Lpgvector_rx_b200 -lpgvrx_b200
//       -Wl,-rpath,$PWD/pgvector_rx_b200 -o pgvrx_bench
//   ./pgvrx_bench --n 1000000 --dim 128 --m 16 --efc 64 --ef 40 --k 10 --nq 10000 --steps 50 --warmup 5 --gpus 1
//
// What it does, in the reference's terms:
//   CREATE INDEX    pgvrx_index_build on GPU 0 (levels drawn with build.rs:373-377's formula through
//                   pgvrx_level_from_uniform; rows are synthetic N(0,1), its own counter-based generator)
//   replicas        pgvrx_index_export_graph -> pgvrx_index_load_graph on GPUs 1..G-1 (the build does not shard)
//   index scans     one host thread per GPU (the reference's one-backend-one-thread model, build.rs:380-385), each
//                   calling the blocking pgvrx_search_batch on its own nq-query batches -- host buffers in, host
//                   buffers out, i.e. the e2e figure; the library's CUDA-event kernel time is reported beside it
//   ground truth    pgvrx_flat_topk (exact kNN) on GPU 0 for recall@k
// It prints ONE JSON line.  Without a usable device every compute call fails with a status code and a message, and so
// does this program (exit code 3): there is no CPU fallback to measure.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "pgvrx_b200.h"

namespace {

struct Args {
    int64_t n = 100000, nq = 10000;
    int dim = 128, m = 16, efc = 64, ef = 40, k = 10, steps = 20, warmup = 3, gpus = 1, threads = 0, batches = 4;
    bool oversubscribe = false;      // place replica g on device g % visible devices (exercises the replica path on one GPU)
    int64_t recall_queries = 1000;
};

bool parse(int argc, char **argv, Args &a) {
    for (int i = 1; i < argc; i++) {
        const std::string f = argv[i];
        auto val = [&](int64_t &dst) { if (i + 1 >= argc) return false; dst = std::strtoll(argv[++i], nullptr, 10); return true; };
        auto vali = [&](int &dst) { int64_t v; if (!val(v)) return false; dst = (int)v; return true; };
        bool ok = true;
        if (f == "--n") ok = val(a.n);
        else if (f == "--nq") ok = val(a.nq);
        else if (f == "--dim") ok = vali(a.dim);
        else if (f == "--m") ok = vali(a.m);
        else if (f == "--efc") ok = vali(a.efc);
        else if (f == "--ef") ok = vali(a.ef);
        else if (f == "--k") ok = vali(a.k);
        else if (f == "--steps") ok = vali(a.steps);
        else if (f == "--warmup") ok = vali(a.warmup);
        else if (f == "--gpus") ok = vali(a.gpus);
        else if (f == "--threads") ok = vali(a.threads);
        else if (f == "--batches") ok = vali(a.batches);
        else if (f == "--recall-queries") ok = val(a.recall_queries);
        else if (f == "--oversubscribe") a.oversubscribe = true;
        else ok = false;
        if (!ok) {
            std::fprintf(stderr, "usage: pgvrx_bench [--n N] [--dim D] [--m M] [--efc E] [--ef EF] [--k K] [--nq Q] [--steps S] "
                                 "[--warmup W] [--gpus G] [--threads T] [--batches B] [--recall-queries R] [--oversubscribe]\n");
            return false;
        }
    }
    return a.n > 0 && a.nq > 0 && a.dim > 0 && a.steps > 0 && a.warmup >= 0 && a.gpus > 0 && a.batches > 0 && a.k > 0;
}

// counter-based generator: element (stream, i) does not depend on how the work is split across threads
inline uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
inline double uniform01(uint64_t stream, uint64_t i) {          // in (0, 1)
    return ((double)(mix64(stream * 0x100000001b3ull + i) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
void fill_gaussian(float *dst, int64_t count, uint64_t stream, int nthreads) {
    std::vector<std::thread> ts;
    const int64_t pairs = (count + 1) / 2, per = (pairs + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++)
        ts.emplace_back([=]() {
            const int64_t lo = t * per, hi = std::min(pairs, lo + per);
            for (int64_t p = lo; p < hi; p++) {                 // Box-Muller, two values per pair of draws
                const double u = uniform01(stream, 2 * (uint64_t)p), v = uniform01(stream, 2 * (uint64_t)p + 1);
                const double r = std::sqrt(-2.0 * std::log(u)), th = 6.283185307179586 * v;
                dst[2 * p] = (float)(r * std::cos(th));
                if (2 * p + 1 < count) dst[2 * p + 1] = (float)(r * std::sin(th));
            }
        });
    for (auto &t : ts) t.join();
}

struct Barrier {                                                // start / stop line of the timed region
    std::mutex mu;
    std::condition_variable cv;
    int count, waiting = 0, generation = 0;
    explicit Barrier(int n) : count(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        const int gen = generation;
        if (++waiting == count) { waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != generation; });
    }
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int fail(const pgvrx_index *ix, const char *what, int rc) {
    std::fprintf(stderr, "pgvrx_bench: %s failed (status %d): %s\n", what, rc, pgvrx_last_error(ix));
    return 3;
}

}  // namespace

int main(int argc, char **argv) {
    Args a;
    if (!parse(argc, argv, a)) return 2;
    if (pgvrx_abi_version() != PGVRX_ABI_VERSION) { std::fprintf(stderr, "pgvrx_bench: ABI version mismatch\n"); return 2; }
    const int ndev = pgvrx_device_count();
    if (ndev < 1 || (ndev < a.gpus && !a.oversubscribe)) {
        std::fprintf(stderr, "pgvrx_bench: %d CUDA device(s) visible, %d requested; the library has no CPU path (%s)\n",
                     ndev < 0 ? 0 : ndev, a.gpus, pgvrx_last_error(nullptr));
        return 3;
    }
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    const int nthreads = a.threads > 0 ? a.threads : hw;
    const size_t row_bytes = pgvrx_row_bytes(PGVRX_ELEM_F32, a.dim);

    // ---- synthetic table, levels, query batches ----
    std::vector<float> rows((size_t)a.n * a.dim);
    fill_gaussian(rows.data(), (int64_t)rows.size(), 2, nthreads);
    std::vector<int32_t> levels((size_t)a.n);
    for (int64_t i = 0; i < a.n; i++) levels[(size_t)i] = pgvrx_level_from_uniform(uniform01(4, (uint64_t)i), a.m);
    std::vector<std::vector<float>> batches((size_t)a.gpus * a.batches);
    for (size_t b = 0; b < batches.size(); b++) {
        batches[b].resize((size_t)a.nq * a.dim);
        fill_gaussian(batches[b].data(), (int64_t)batches[b].size(), 1000 + b, nthreads);
    }

    // ---- CREATE INDEX on GPU 0 ----
    pgvrx_index_desc d;
    std::memset(&d, 0, sizeof d);
    d.elem_type = PGVRX_ELEM_F32; d.metric = PGVRX_METRIC_L2SQ; d.dim = a.dim; d.m = a.m; d.ef_construction = a.efc;
    d.device = 0; d.capacity = a.n;
    std::vector<pgvrx_index *> ix((size_t)a.gpus, nullptr);
    int rc = pgvrx_index_create(&d, &ix[0]);
    if (rc) return fail(nullptr, "pgvrx_index_create", rc);
    pgvrx_build_stats bs;
    const double tb0 = now_s();
    rc = pgvrx_index_build(ix[0], rows.data(), a.n, levels.data(), nullptr, a.threads, &bs);
    if (rc) return fail(ix[0], "pgvrx_index_build", rc);
    const double build_s = now_s() - tb0;

    // ---- replicas on the other GPUs ----
    if (a.gpus > 1) {
        int64_t n = 0, nu = 0, entry = -1;
        int32_t entry_level = 0;
        rc = pgvrx_index_graph_info(ix[0], &n, &nu, &entry, &entry_level);
        if (rc) return fail(ix[0], "pgvrx_index_graph_info", rc);
        std::vector<uint8_t> level((size_t)n);
        std::vector<int32_t> adj0((size_t)n * 2 * a.m), upper_base((size_t)n), adj_up((size_t)std::max<int64_t>(nu, 1) * a.m);
        std::vector<uint8_t> grows((size_t)n * row_bytes);     // element rows (duplicates merged: n may be < a.n)
        rc = pgvrx_index_export_graph(ix[0], grows.data(), level.data(), adj0.data(), upper_base.data(), adj_up.data(), nullptr, nullptr);
        if (rc) return fail(ix[0], "pgvrx_index_export_graph", rc);
        for (int g = 1; g < a.gpus; g++) {
            pgvrx_index_desc dg = d;
            dg.device = g % ndev; dg.capacity = n;
            rc = pgvrx_index_create(&dg, &ix[(size_t)g]);
            if (rc) return fail(nullptr, "pgvrx_index_create (replica)", rc);
            rc = pgvrx_index_load_graph(ix[(size_t)g], grows.data(), n, level.data(), adj0.data(), upper_base.data(), adj_up.data(), nu, entry);
            if (rc) return fail(ix[(size_t)g], "pgvrx_index_load_graph", rc);
        }
    }

    // ---- index scans: one host thread per GPU, blocking calls on host buffers ----
    struct PerGpu {
        std::vector<int32_t> ids, counts;
        std::vector<float> dists;
        double kernel_ms = 0.0, wall_s = 0.0;
        uint64_t n_dist = 0, n_expand0 = 0, n_expand_upper = 0, launches = 0;
        int rc = 0;
    };
    std::vector<PerGpu> per((size_t)a.gpus);
    Barrier bar(a.gpus);
    std::atomic<int> failed{0};
    std::vector<std::thread> workers;
    for (int g = 0; g < a.gpus; g++)
        workers.emplace_back([&, g]() {
            PerGpu &p = per[(size_t)g];
            p.ids.resize((size_t)a.nq * a.k); p.dists.resize((size_t)a.nq * a.k); p.counts.resize((size_t)a.nq);
            auto step = [&](int s, pgvrx_search_stats *st) {
                const std::vector<float> &q = batches[(size_t)g * a.batches + (size_t)(s % a.batches)];
                return pgvrx_search_batch(ix[(size_t)g], q.data(), a.nq, a.ef, a.k, p.ids.data(), p.dists.data(), p.counts.data(), st);
            };
            for (int s = 0; s < a.warmup && !p.rc; s++) p.rc = step(s, nullptr);
            if (p.rc) failed++;
            bar.wait();
            const double t0 = now_s();
            for (int s = 0; s < a.steps && !p.rc && !failed.load(); s++) {
                pgvrx_search_stats st;
                p.rc = step(a.warmup + s, &st);
                p.kernel_ms += st.kernel_ms; p.n_dist += st.n_dist; p.n_expand0 += st.n_expand0;
                p.n_expand_upper += st.n_expand_upper; p.launches += st.n_launches;
            }
            p.wall_s = now_s() - t0;
            if (p.rc) failed++;
            bar.wait();
        });
    for (auto &w : workers) w.join();
    for (int g = 0; g < a.gpus; g++)
        if (per[(size_t)g].rc) return fail(ix[(size_t)g], "pgvrx_search_batch", per[(size_t)g].rc);
    double wall_max = 0.0, kernel_max = 0.0;
    uint64_t n_dist = 0, n_e0 = 0, n_eu = 0, launches = 0;
    for (const PerGpu &p : per) {
        wall_max = std::max(wall_max, p.wall_s); kernel_max = std::max(kernel_max, p.kernel_ms);
        n_dist += p.n_dist; n_e0 += p.n_expand0; n_eu += p.n_expand_upper; launches += p.launches;
    }
    const double total_q = (double)a.gpus * a.nq * a.steps;
    const double gathered = (double)n_dist * row_bytes + (double)n_e0 * 2 * a.m * 4 + (double)n_eu * a.m * 4;

    // ---- recall@k against the exact scan, first R queries of GPU 0's first batch ----
    const int64_t R = std::min<int64_t>(a.recall_queries, a.nq);
    double recall = -1.0;
    int64_t full_counts = 0;
    if (R > 0) {
        std::vector<int32_t> sid((size_t)R * a.k), scnt((size_t)R), fid((size_t)R * a.k);
        std::vector<float> sd((size_t)R * a.k), fd((size_t)R * a.k);
        rc = pgvrx_search_batch(ix[0], batches[0].data(), R, a.ef, a.k, sid.data(), sd.data(), scnt.data(), nullptr);
        if (rc) return fail(ix[0], "pgvrx_search_batch (recall)", rc);
        rc = pgvrx_flat_topk(ix[0], batches[0].data(), R, a.k, fid.data(), fd.data(), nullptr);
        if (rc) return fail(ix[0], "pgvrx_flat_topk", rc);
        int64_t hits = 0;
        for (int64_t q = 0; q < R; q++) {
            full_counts += scnt[(size_t)q] == std::min<int64_t>(a.k, a.n);
            for (int i = 0; i < scnt[(size_t)q]; i++)
                for (int j = 0; j < a.k; j++)
                    if (sid[(size_t)q * a.k + i] == fid[(size_t)q * a.k + j]) { hits++; break; }
            for (int i = 1; i < scnt[(size_t)q]; i++)               // rows come back nearest first (scan.rs:441-447 reversed)
                if (sd[(size_t)q * a.k + i] < sd[(size_t)q * a.k + i - 1]) {
                    std::fprintf(stderr, "pgvrx_bench: query %lld is not sorted by distance\n", (long long)q);
                    return 4;
                }
        }
        recall = (double)hits / ((double)R * std::min<int64_t>(a.k, a.n));
    }
    // ---- the same GPUs behind ONE handle (pgvrx_group_*): (a) exact flat kNN over row shards, per-shard lists merged by
    //      one kernel on the lead device that reads the peers' lists over NVLink; (b) one caller batch split by queries
    //      over the replicas (strong scaling of a single call) ----
    double grp_flat_qps = 0.0, grp_flat_ms = 0.0, grp_merge_ms = 0.0, grp_scan_qps = 0.0, grp_scan_ms = 0.0;
    int grp_peer = 0, grp_flat_ok = -1, grp_scan_ok = -1;
    {
        std::vector<int32_t> devs((size_t)a.gpus);
        for (int g = 0; g < a.gpus; g++) devs[(size_t)g] = g % ndev;
        int64_t n = 0, nu = 0, entry = -1;
        int32_t entry_level = 0;
        rc = pgvrx_index_graph_info(ix[0], &n, &nu, &entry, &entry_level);
        if (rc) return fail(ix[0], "pgvrx_index_graph_info", rc);
        std::vector<uint8_t> level((size_t)n);
        std::vector<int32_t> adj0((size_t)n * 2 * a.m), upper_base((size_t)n), adj_up((size_t)std::max<int64_t>(nu, 1) * a.m);
        std::vector<uint8_t> grows((size_t)n * row_bytes);
        rc = pgvrx_index_export_graph(ix[0], grows.data(), level.data(), adj0.data(), upper_base.data(), adj_up.data(), nullptr, nullptr);
        if (rc) return fail(ix[0], "pgvrx_index_export_graph", rc);
        pgvrx_index_desc dg = d;
        dg.capacity = n;
        pgvrx_group *grp = nullptr;
        rc = pgvrx_group_create(&dg, devs.data(), a.gpus, &grp);
        if (rc) { std::fprintf(stderr, "pgvrx_bench: pgvrx_group_create failed (status %d): %s\n", rc, pgvrx_group_last_error(nullptr)); return 3; }
        auto gfail = [&](const char *what, int code) {
            std::fprintf(stderr, "pgvrx_bench: %s failed (status %d): %s\n", what, code, pgvrx_group_last_error(grp));
            return 3;
        };
        const std::vector<float> &q = batches[0];
        std::vector<int32_t> gid((size_t)a.nq * a.k), gcnt((size_t)a.nq), wid((size_t)a.nq * a.k), wcnt((size_t)a.nq);
        std::vector<float> gd((size_t)a.nq * a.k), wd((size_t)a.nq * a.k);
        const int reps = std::max(3, std::min(a.steps, 10));
        pgvrx_group_stats gs;
        // (a) sharded flat
        rc = pgvrx_group_load_rows(grp, grows.data(), n);
        if (rc) return gfail("pgvrx_group_load_rows", rc);
        for (int r = 0; r < 2 && !rc; r++) rc = pgvrx_group_flat_topk(grp, q.data(), a.nq, a.k, gid.data(), gd.data(), &gs);
        if (rc) return gfail("pgvrx_group_flat_topk", rc);
        double t0 = now_s();
        for (int r = 0; r < reps && !rc; r++) { rc = pgvrx_group_flat_topk(grp, q.data(), a.nq, a.k, gid.data(), gd.data(), &gs); grp_merge_ms += gs.merge_kernel_ms; }
        if (rc) return gfail("pgvrx_group_flat_topk", rc);
        grp_flat_ms = (now_s() - t0) * 1e3 / reps;
        grp_flat_qps = a.nq / (grp_flat_ms / 1e3);
        grp_merge_ms /= reps;
        grp_peer = gs.n_peer_direct;
        const int64_t RC = std::min<int64_t>(R > 0 ? R : 256, a.nq);
        rc = pgvrx_flat_topk(ix[0], q.data(), RC, a.k, wid.data(), wd.data(), nullptr);
        if (rc) return fail(ix[0], "pgvrx_flat_topk (group check)", rc);
        grp_flat_ok = std::memcmp(gid.data(), wid.data(), (size_t)RC * a.k * 4) == 0 && std::memcmp(gd.data(), wd.data(), (size_t)RC * a.k * 4) == 0;
        // (b) query-sharded scan of one batch
        rc = pgvrx_group_load_graph(grp, grows.data(), n, level.data(), adj0.data(), upper_base.data(), adj_up.data(), nu, entry);
        if (rc) return gfail("pgvrx_group_load_graph", rc);
        for (int r = 0; r < 2 && !rc; r++) rc = pgvrx_group_search_batch(grp, q.data(), a.nq, a.ef, a.k, gid.data(), gd.data(), gcnt.data(), nullptr);
        if (rc) return gfail("pgvrx_group_search_batch", rc);
        t0 = now_s();
        for (int r = 0; r < reps && !rc; r++) rc = pgvrx_group_search_batch(grp, q.data(), a.nq, a.ef, a.k, gid.data(), gd.data(), gcnt.data(), nullptr);
        if (rc) return gfail("pgvrx_group_search_batch", rc);
        grp_scan_ms = (now_s() - t0) * 1e3 / reps;
        grp_scan_qps = a.nq / (grp_scan_ms / 1e3);
        rc = pgvrx_search_batch(ix[0], q.data(), a.nq, a.ef, a.k, wid.data(), wd.data(), wcnt.data(), nullptr);
        if (rc) return fail(ix[0], "pgvrx_search_batch (group check)", rc);
        grp_scan_ok = std::memcmp(gid.data(), wid.data(), gid.size() * 4) == 0 && std::memcmp(gcnt.data(), wcnt.data(), gcnt.size() * 4) == 0;
        pgvrx_group_destroy(grp);
    }
    for (pgvrx_index *p : ix) pgvrx_index_destroy(p);

    std::printf("{\"binary\": \"pgvrx_bench\", \"metric\": \"QPS@recall@10 and build vectors/s\", \"value\": %.1f, \"unit\": \"queries/s\", "
                "\"n_gpus\": %d, \"n_devices\": %d, \"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.4f, \"kernel_qps\": %.1f, \"kernel_ms_per_step\": %.4f, "
                "\"gpu_launches\": %llu, \"recall_at_k\": %.5f, \"recall_queries\": %lld, \"full_result_sets\": %lld, "
                "\"gathered_gb_per_s_kernel\": %.1f, \"n_dist\": %llu, \"n_expand0\": %llu, \"n_expand_upper\": %llu, "
                "\"build\": {\"vectors_per_s\": %.1f, \"seconds\": %.3f, \"rounds\": %llu, \"conflicts\": %llu, \"duplicates\": %llu, \"host_threads\": %d}, "
                "\"group\": {\"members\": %d, \"flat_sharded_qps\": %.1f, \"flat_sharded_ms_per_batch\": %.4f, \"merge_kernel_ms\": %.4f, "
                "\"lists_read_in_place\": %d, \"flat_equals_single_table\": %d, \"scan_one_batch_qps\": %.1f, \"scan_one_batch_ms\": %.4f, "
                "\"scan_equals_single_device\": %d}, "
                "\"config\": {\"workload\": \"%lldx%d fp32 vector_l2_ops m=%d ef_construction=%d ef_search=%d k=%d, %lld-query batches per GPU\", "
                "\"path\": \"blocking pgvrx_search_batch on pageable host buffers, one host thread per GPU, graph replicated\"}}\n",
                total_q / wall_max, a.gpus, std::min(a.gpus, ndev), a.steps, a.warmup, wall_max * 1e3 / a.steps, total_q / (kernel_max / 1e3),
                kernel_max / a.steps, (unsigned long long)launches, recall, (long long)R, (long long)full_counts,
                kernel_max > 0 ? gathered / a.gpus / (kernel_max / 1e3) / 1e9 : 0.0, (unsigned long long)n_dist, (unsigned long long)n_e0,
                (unsigned long long)n_eu, (double)a.n / build_s, build_s, (unsigned long long)bs.n_rounds,
                (unsigned long long)bs.n_conflicts, (unsigned long long)bs.n_duplicates, bs.host_threads,
                a.gpus, grp_flat_qps, grp_flat_ms, grp_merge_ms, grp_peer, grp_flat_ok, grp_scan_qps, grp_scan_ms, grp_scan_ok,
                (long long)a.n, a.dim, a.m, a.efc, a.ef, a.k, (long long)a.nq);
    return 0;
}
