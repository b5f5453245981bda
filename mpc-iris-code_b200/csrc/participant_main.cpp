// iris_participant -- wire-compatible stand-in for the reference's `participant` subcommand
// (src/main.rs:384-452) on top of the C ABI, with the share database resident in HBM.
//
// Protocol (raw native-endian Pod bytes, no framing; src/main.rs:418-443):
//   request : one 3 200-byte Template {pattern: Bits, mask: Bits}           (src/template.rs:11-29)
//   reply   : [u16;31] per database row (62 bytes, rotation -15..=15), streamed back; EOF terminates the
//             reply.  The reference produces it in batches of 20 000 rows through an mpsc(4) channel
//             (src/main.rs:423-443); the batch boundaries are not visible on the wire.
// One request at a time, like the reference.  An unmodified reference coordinator / benchmark
// (src/main.rs:486-504, 645-686) can connect to this process.
//
// Same two-stage shape as the reference: a worker thread produces, the connection's loop streams.
//   one GPU      the worker scans batch after batch into a ring of three page-locked buffers (the channel) and the
//                main thread writes finished batches to the socket.  The ring's back-pressure keeps the producer one or
//                two batches ahead of the socket, so the bytes write() reads are still in the CPU's cache: over
//                loopback the single TCP stream is the limit (7.4 ms per 1 M rows = 8.4 GB/s on the wire; the scan
//                alone is 3.6 ms).  [One whole-range call into a 62 MB array frees the GPU after 3.9 ms with the first
//                byte out after 1 ms, but write() then reads cold memory and the request takes 10 ms.]
//   --devices    the share file is row-sharded over several GPUs (iris_cluster_*): ONE library call per request scans
//                every block at the same time into one page-locked array and reports the rows that have landed; the
//                main thread writes the completed prefix.  The wire is the limit by a wide margin here.
//
//   iris_participant --input mpc.share-0 [--bind 127.0.0.1:1234] [--device 0 | --devices 0,1,2,3 | --devices 0-7]
//                    [--batch-size 20000] [--max-requests N] [--synthetic ROWS --seed S]
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/iris_b200.h"

static void die(const char* what) {
    fprintf(stderr, "iris_participant: %s: %s\n", what, iris_last_error());
    exit(1);
}

static bool read_exact(int fd, void* buf, size_t n) {
    uint8_t* p = static_cast<uint8_t*>(buf);
    while (n) {
        ssize_t r = read(fd, p, n);
        if (r == 0) return false;
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += r;
        n -= (size_t)r;
    }
    return true;
}

static bool write_all(int fd, const void* buf, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(buf);
    while (n) {
        ssize_t r = write(fd, p, n);
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += r;
        n -= (size_t)r;
    }
    return true;
}

int main(int argc, char** argv) {
    std::string input, bind_addr = "127.0.0.1:1234";   // reference default, src/main.rs:124
    std::vector<int> devices{0};
    uint64_t batch = 20000, synthetic = 0, seed = 0x1715C0DE;   // the reference's chunk, src/main.rs:428
    long max_requests = -1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* {
            if (i + 1 >= argc) {
                fprintf(stderr, "missing value for %s\n", a.c_str());
                exit(2);
            }
            return argv[++i];
        };
        if (a == "--input") input = next();
        else if (a == "--bind") bind_addr = next();
        else if (a == "--device") devices = {atoi(next())};
        else if (a == "--devices") {                      // "0,1,2" or "0-7"
            devices.clear();
            std::string v = next();
            size_t pos = 0;
            while (pos <= v.size()) {
                size_t comma = v.find(',', pos);
                if (comma == std::string::npos) comma = v.size();
                const std::string item = v.substr(pos, comma - pos);
                const size_t dash = item.find('-');
                if (item.empty()) {
                } else if (dash == std::string::npos) {
                    devices.push_back(atoi(item.c_str()));
                } else {
                    for (int d = atoi(item.substr(0, dash).c_str()); d <= atoi(item.substr(dash + 1).c_str()); ++d) devices.push_back(d);
                }
                pos = comma + 1;
            }
            if (devices.empty()) {
                fprintf(stderr, "bad --devices %s\n", v.c_str());
                return 2;
            }
        }
        else if (a == "--batch-size") batch = strtoull(next(), nullptr, 10);
        else if (a == "--max-requests") max_requests = atol(next());
        else if (a == "--synthetic") synthetic = strtoull(next(), nullptr, 10);
        else if (a == "--seed") seed = strtoull(next(), nullptr, 0);
        else {
            fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    if (input.empty() && !synthetic) {
        fprintf(stderr, "usage: iris_participant --input <share file> | --synthetic <rows> [--bind host:port]\n");
        return 2;
    }
    if (batch == 0) batch = 20000;
    signal(SIGPIPE, SIG_IGN);

    // "Opened share ... with N encrypted patterns" (src/main.rs:386-400): here the file goes to HBM once.
    uint64_t rows = synthetic;
    if (!synthetic) {
        struct stat st;
        if (stat(input.c_str(), &st) != 0) {
            fprintf(stderr, "Failed to open share at %s\n", input.c_str());
            return 1;
        }
        if (st.st_size % (IRIS_BITS * 2)) {
            fprintf(stderr, "Share file %s invalid.\n", input.c_str());
            return 1;
        }
        rows = (uint64_t)st.st_size / (IRIS_BITS * 2);
    }
    // one handle for the whole file: contiguous row blocks in the HBM of each GPU, loaded by all GPUs at the same time
    iris_cluster* cluster = nullptr;
    if (iris_cluster_create(devices.data(), (uint32_t)devices.size(), rows ? rows : 1, IRIS_DB_SHARES, &cluster)) die("iris_db_create");
    if (synthetic) {
        if (iris_cluster_generate(cluster, seed, 0, 0, 0, rows)) die("iris_cluster_generate");
    } else if (rows) {
        if (iris_cluster_load_files(cluster, input.c_str(), nullptr)) die("iris_cluster_load_files");
    }
    const int device = devices[0];
    iris_db* db = nullptr;                                   // the single shard of a one-GPU participant
    if (iris_cluster_shard(cluster, 0, &db, nullptr, nullptr, nullptr)) die("iris_cluster_shard");
    fprintf(stderr, "Opened share with %llu encrypted patterns (resident in HBM on %zu GPU%s)\n", (unsigned long long)rows,
            devices.size(), devices.size() == 1 ? "" : "s");

    const size_t colon = bind_addr.rfind(':');
    if (colon == std::string::npos) {
        fprintf(stderr, "bad --bind %s\n", bind_addr.c_str());
        return 2;
    }
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)atoi(bind_addr.c_str() + colon + 1));
    if (inet_pton(AF_INET, bind_addr.substr(0, colon).c_str(), &addr.sin_addr) != 1) {
        fprintf(stderr, "bad --bind %s\n", bind_addr.c_str());
        return 2;
    }
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    if (bind(ls, reinterpret_cast<sockaddr*>(&addr), sizeof addr) != 0 || listen(ls, 16) != 0) {
        fprintf(stderr, "Could not bind to socket %s: %s\n", bind_addr.c_str(), strerror(errno));
        return 1;
    }
    socklen_t alen = sizeof addr;
    getsockname(ls, reinterpret_cast<sockaddr*>(&addr), &alen);
    fprintf(stderr, "Listening on %s:%d\n", bind_addr.substr(0, colon).c_str(), (int)ntohs(addr.sin_port));
    fflush(stderr);

    constexpr uint64_t kRing = 3;                            // batches in flight between the scan and the socket
    uint16_t* ring[kRing];
    for (uint16_t*& slot : ring)
        if (iris_host_alloc(batch * IRIS_ROTATIONS * sizeof(uint16_t), reinterpret_cast<void**>(&slot))) die("iris_host_alloc");
    uint16_t* whole = nullptr;                               // several GPUs: the complete reply, written by all of them
    if (devices.size() > 1 && iris_host_alloc((rows ? rows : 1) * IRIS_ROTATIONS * sizeof(uint16_t), reinterpret_cast<void**>(&whole)))
        die("iris_host_alloc");
    uint32_t n_shards = 0;
    if (iris_cluster_len(cluster, &n_shards, nullptr, nullptr)) die("iris_cluster_len");
    std::vector<uint64_t> shard_begin(n_shards), shard_end(n_shards);
    for (uint32_t i = 0; i < n_shards; ++i)
        if (iris_cluster_shard(cluster, i, nullptr, nullptr, &shard_begin[i], &shard_end[i])) die("iris_cluster_shard");
    struct Shared {
        std::mutex mu;
        std::condition_variable cv;
        std::vector<uint64_t> done;          // per shard: cluster rows [shard_begin, done) are in host memory
        const std::vector<uint64_t>* begin = nullptr;
        bool finished = false;
        int rc = 0;
        std::string error;
    } sh;
    sh.begin = &shard_begin;
    // called by the library's per-GPU threads as blocks of rows land in `whole`
    auto on_progress = [](void* user, uint64_t, uint64_t e) {
        Shared* s = static_cast<Shared*>(user);
        {
            std::lock_guard<std::mutex> lk(s->mu);
            size_t i = s->begin->size();
            while (i-- > 0)
                if (e > (*s->begin)[i]) break;               // the shard whose block contains row e - 1
            s->done[i] = e;
        }
        s->cv.notify_all();
    };
    const uint64_t n_batches = (rows + batch - 1) / batch;
    uint64_t tmpl[2 * IRIS_LIMBS];
    for (long served = 0; max_requests < 0 || served < max_requests; ++served) {
        int fd = accept(ls, nullptr, nullptr);
        if (fd < 0) {
            if (errno == EINTR) {
                --served;
                continue;
            }
            break;
        }
        setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
        if (!read_exact(fd, tmpl, sizeof tmpl)) {           // stream.read_exact(bytes_of_mut(&mut template))
            close(fd);
            continue;
        }
        fprintf(stderr, "Request received.\n");
        if (whole) {
            // several GPUs: encode(&template), DistanceEngine::new, batch_process over every block at the same time
            // (src/main.rs:427-430) in ONE library call; the rows that have landed are streamed in row order
            {
                std::lock_guard<std::mutex> lk(sh.mu);
                sh.done = shard_begin;
                sh.finished = false;
                sh.rc = 0;
            }
            std::thread producer([&] {
                const int rc = rows ? iris_cluster_match_template_streamed(cluster, tmpl, tmpl + IRIS_LIMBS, whole, nullptr,
                                                                           on_progress, &sh)
                                    : 0;
                {
                    std::lock_guard<std::mutex> lk(sh.mu);
                    if (rc) sh.error = iris_last_error();    // thread-local: capture it on this thread
                    sh.rc = rc;
                    sh.finished = true;
                }
                sh.cv.notify_all();
            });
            bool sent_all = true;
            uint64_t sent = 0;
            while (sent_all && sent < rows) {
                uint64_t upto = sent;
                {
                    std::unique_lock<std::mutex> lk(sh.mu);
                    auto prefix = [&] {
                        uint64_t p = 0;
                        for (uint32_t i = 0; i < n_shards; ++i) {
                            p = sh.done[i];
                            if (sh.done[i] < shard_end[i]) break;
                        }
                        return p;
                    };
                    sh.cv.wait(lk, [&] { return prefix() > sent || (sh.finished && sh.rc); });
                    if (sh.finished && sh.rc) break;
                    upto = prefix();
                }
                sent_all = write_all(fd, whole + sent * IRIS_ROTATIONS, (upto - sent) * IRIS_ROTATIONS * sizeof(uint16_t));
                sent = upto;
            }
            producer.join();
            close(fd);
            if (sh.rc) {
                fprintf(stderr, "iris_participant: batch_process: %s\n", sh.error.c_str());
                return 1;
            }
            fprintf(stderr, sent_all ? "Reply sent.\n" : "Peer went away.\n");
            continue;
        }
        iris_distance_engine* engine = nullptr;
        if (iris_distance_engine_new_from_template(device, tmpl, tmpl + IRIS_LIMBS, &engine)) die("engine");
        std::mutex mu;
        std::condition_variable cv;
        uint64_t produced = 0, consumed = 0;                 // batches scanned / batches written
        bool stop = false;
        std::string worker_error;
        // spawn_blocking worker of src/main.rs:425-434: for chunk in patterns.chunks(..) { batch_process; send }
        std::thread worker([&] {
            for (uint64_t i = 0; i < n_batches; ++i) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || i - consumed < kRing; });
                    if (stop) return;
                }
                const uint64_t b = i * batch, e = b + batch < rows ? b + batch : rows;
                const int rc = iris_distance_engine_batch_process_resident(engine, ring[i % kRing], e - b, db, b, e);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (rc) {
                        worker_error = iris_last_error();    // thread-local: capture it on this thread
                        stop = true;
                    } else {
                        produced = i + 1;
                    }
                }
                cv.notify_all();
                if (rc) return;
            }
        });
        // "Stream output" loop of src/main.rs:437-443
        bool ok = true;
        for (uint64_t i = 0; ok && i < n_batches; ++i) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || produced > i; });
                if (produced <= i) break;                    // the worker failed
            }
            const uint64_t b = i * batch, e = b + batch < rows ? b + batch : rows;
            ok = write_all(fd, ring[i % kRing], (e - b) * IRIS_ROTATIONS * sizeof(uint16_t));
            {
                std::lock_guard<std::mutex> lk(mu);
                consumed = i + 1;
            }
            cv.notify_all();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        worker.join();
        iris_distance_engine_free(engine);
        close(fd);                                           // EOF = end of results
        if (!worker_error.empty()) {
            fprintf(stderr, "iris_participant: batch_process: %s\n", worker_error.c_str());
            return 1;
        }
        fprintf(stderr, ok ? "Reply sent.\n" : "Peer went away.\n");
    }
    for (uint16_t* slot : ring) iris_host_free(slot);
    iris_host_free(whole);
    close(ls);
    iris_cluster_destroy(cluster);
    return 0;
}
