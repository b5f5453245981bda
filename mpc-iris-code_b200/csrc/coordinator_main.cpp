// iris_coordinator -- wire-compatible stand-in for the query loop of the reference's `coordinator` subcommand
// (src/main.rs:453-640) on top of the C ABI, with the mask database resident in HBM.
//
// Per request (src/main.rs:476-637):
//   1. connect to every participant and send the 3 200-byte Template           (src/main.rs:486-504)
//   2. denominators of the query mask against the local masks                  (src/main.rs:507-519) -- here ONE
//      device scan over the resident masks instead of 20 000-row CPU batches
//   3. read [u16;31] rows from every participant in batches of 20 000, cut every batch to the shortest
//      prefix, stop at the first empty batch                                   (src/main.rs:522-578)
//   4. numerator = wrapping sum of the shares, decode_distance, running minimum with `<`
//                                                                              (src/main.rs:597-621) -- on the device
//   5. report "Found closest entry at {index} out of {rows} at distance {d}."  (src/main.rs:634-636)
// Unmodified reference participants (src/main.rs:384-452) can serve this process, and iris_participant can serve
// an unmodified reference coordinator.
//
// The reference draws a random Template per request (src/main.rs:479); this front-end takes the queries from a
// file of 3 200-byte Templates (--queries) or draws them from a seeded generator (--seed), and stops after
// --requests queries instead of looping forever.  Each result also goes to stdout as
// "<min_index> <rows> <min_distance %.17g>".
//
//   iris_coordinator --masks mpc.masks [--device 0] [--queries FILE | --seed S] [--requests N]
//                    [--batch-rows 20000] HOST:PORT...
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cinttypes>
#include <cmath>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/iris_b200.h"

namespace {

constexpr size_t kRowBytes = IRIS_ROTATIONS * sizeof(uint16_t);          // size_of::<[u16; 31]>()
constexpr size_t kTemplateBytes = 2 * IRIS_LIMBS * sizeof(uint64_t);     // src/template.rs:11-29

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "iris_coordinator: %s: %s\n", what, iris_last_error());
    exit(1);
}

bool write_all(int fd, const void* buf, size_t n) {
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

// Fill `buf` as far as the stream allows; returns the bytes read (short only at end of stream).
size_t read_upto(int fd, void* buf, size_t n) {
    uint8_t* p = static_cast<uint8_t*>(buf);
    size_t got = 0;
    while (got < n) {
        ssize_t r = read(fd, p + got, n - got);
        if (r == 0) break;
        if (r < 0) {
            if (errno == EINTR) continue;
            break;
        }
        got += (size_t)r;
    }
    return got;
}

int connect_to(const std::string& address) {
    const size_t colon = address.rfind(':');
    if (colon == std::string::npos) return -1;
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)atoi(address.c_str() + colon + 1));
    if (inet_pton(AF_INET, address.substr(0, colon).c_str(), &addr.sin_addr) != 1) return -1;
    int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    if (connect(fd, reinterpret_cast<sockaddr*>(&addr), sizeof addr) != 0) {
        close(fd);
        return -1;
    }
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    return fd;
}

uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace

int main(int argc, char** argv) {
    std::string masks_path = "mpc.masks", queries_path;                  // default: src/main.rs:131
    std::vector<std::string> participants;
    int device = 0;
    uint64_t seed = 0x1715C0DE;
    uint64_t batch_rows = 20000;                                         // BATCH_SIZE, src/main.rs:473
    long requests = 1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* {
            if (i + 1 >= argc) {
                fprintf(stderr, "missing value for %s\n", a.c_str());
                exit(2);
            }
            return argv[++i];
        };
        if (a == "--masks") masks_path = next();
        else if (a == "--device") device = atoi(next());
        else if (a == "--queries") queries_path = next();
        else if (a == "--seed") seed = strtoull(next(), nullptr, 0);
        else if (a == "--requests") requests = atol(next());
        else if (a == "--batch-rows") batch_rows = strtoull(next(), nullptr, 10);
        else if (a.rfind("--", 0) == 0) {
            fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        } else participants.push_back(a);
    }
    if (participants.empty() || participants.size() > 8) {
        fprintf(stderr, "usage: iris_coordinator --masks <file> [--queries <file> | --seed S] [--requests N] HOST:PORT... (1-8)\n");
        return 2;
    }
    if (batch_rows == 0) batch_rows = 20000;
    signal(SIGPIPE, SIG_IGN);

    // "Opened main ... with N masks" (src/main.rs:455-470): the file goes to HBM once.
    struct stat st;
    if (stat(masks_path.c_str(), &st) != 0) {
        fprintf(stderr, "Failed to open main at %s\n", masks_path.c_str());
        return 1;
    }
    if (st.st_size % (IRIS_LIMBS * 8)) {
        fprintf(stderr, "Main file %s invalid.\n", masks_path.c_str());
        return 1;
    }
    const uint64_t count = (uint64_t)st.st_size / (IRIS_LIMBS * 8);
    iris_db* db = nullptr;
    if (iris_db_create(device, count ? count : 1, IRIS_DB_MASKS, &db)) die("iris_db_create");
    if (count && iris_db_load_masks_file(db, masks_path.c_str(), 0, 0)) die("iris_db_load_masks_file");
    fprintf(stderr, "Opened main with %llu masks (resident in HBM on device %d)\n", (unsigned long long)count, device);

    FILE* qf = nullptr;
    if (!queries_path.empty() && !(qf = fopen(queries_path.c_str(), "rb"))) {
        fprintf(stderr, "Failed to open queries at %s\n", queries_path.c_str());
        return 1;
    }

    const size_t parties = participants.size();
    const size_t batch_bytes = batch_rows * kRowBytes;
    constexpr uint64_t kRing = 3;                                         // batches buffered per participant
    // Denominators stay in HBM; the share batches land in page-locked rings, one reader thread per participant
    // (the reference polls all streams concurrently with try_join_all, src/main.rs:525-560).
    uint16_t* d_denominators = nullptr;
    if (iris_device_alloc(device, (count ? count : 1) * kRowBytes, reinterpret_cast<void**>(&d_denominators)))
        die("iris_device_alloc");
    struct Stream {
        int fd = -1;
        uint16_t* ring[kRing] = {};
        uint64_t rows[kRing] = {};          // whole rows in each buffered batch
        uint64_t produced = 0;              // batches read (guarded by mu)
        bool eof = false;
        std::thread reader;
    };
    std::vector<Stream> streams(parties);
    for (Stream& s : streams)
        for (uint16_t*& slot : s.ring)
            if (iris_host_alloc(batch_bytes, reinterpret_cast<void**>(&slot))) die("iris_host_alloc");
    std::vector<const uint16_t*> share_ptrs(parties);
    uint64_t tmpl[2 * IRIS_LIMBS];
    for (long request = 0; request < requests; ++request) {
        if (qf) {
            if (fread(tmpl, kTemplateBytes, 1, qf) != 1) break;         // out of queries
        } else {
            for (uint64_t& w : tmpl) w = splitmix64(seed);               // thread_rng().gen::<Template>()
        }

        using clock = std::chrono::steady_clock;
        auto ms_since = [](clock::time_point t) { return std::chrono::duration<double, std::milli>(clock::now() - t).count(); };
        const clock::time_point t_request = clock::now();
        double ms_wait = 0, ms_combine = 0;
        for (size_t i = 0; i < parties; ++i) {
            const int fd = connect_to(participants[i]);
            if (fd < 0) {
                fprintf(stderr, "Could not connect to %s\n", participants[i].c_str());
                return 1;
            }
            if (!write_all(fd, tmpl, kTemplateBytes)) {                   // stream.write_all(bytes_of(&query))
                fprintf(stderr, "Could not send the request to %s\n", participants[i].c_str());
                return 1;
            }
            streams[i].fd = fd;
            streams[i].produced = 0;
            streams[i].eof = false;
        }

        std::mutex mu;
        std::condition_variable cv;
        uint64_t consumed = 0;                                            // batches combined
        bool stop = false;
        for (size_t i = 0; i < parties; ++i) {
            Stream& s = streams[i];
            s.reader = std::thread([&, i] {
                for (uint64_t b = 0;; ++b) {
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return stop || b - consumed < kRing; });
                        if (stop) return;
                    }
                    // fill the whole batch unless the stream ends first (src/main.rs:537-556)
                    const size_t got = read_upto(s.fd, s.ring[b % kRing], batch_bytes);
                    if (got < batch_bytes) {
                        fprintf(stderr, "Participant %zu finished.\n", i);
                        if (got % kRowBytes) fprintf(stderr, "Warning: received partial results from %zu.\n", i);
                    }
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        s.rows[b % kRing] = got / kRowBytes;             // whole rows only
                        s.produced = b + 1;
                        s.eof = got < batch_bytes;
                    }
                    cv.notify_all();
                    if (got < batch_bytes) return;
                }
            });
        }

        // MasksEngine::new(&query.mask) + batch_process over all the masks (src/main.rs:507-519), while the
        // participants work
        const double ms_connect = ms_since(t_request);
        const clock::time_point t_den = clock::now();
        iris_masks_engine* engine = nullptr;
        if (iris_masks_engine_new(device, tmpl + IRIS_LIMBS, &engine)) die("iris_masks_engine_new");
        if (count && iris_masks_engine_batch_process_resident(engine, d_denominators, count, db, 0, count))
            die("iris_masks_engine_batch_process_resident");
        if (iris_db_synchronize(db)) die("iris_db_synchronize");
        iris_masks_engine_free(engine);
        const double ms_denominators = ms_since(t_den);

        double min_distance = std::numeric_limits<double>::infinity();
        uint64_t min_index = UINT64_MAX;                                  // usize::MAX
        uint64_t done = 0;
        for (uint64_t b = 0;; ++b) {
            uint64_t batch_size = count - done < batch_rows ? count - done : batch_rows;   // denominators left
            const clock::time_point t_wait = clock::now();
            {
                std::unique_lock<std::mutex> lk(mu);
                for (size_t i = 0; i < parties; ++i) {
                    Stream& s = streams[i];
                    cv.wait(lk, [&] { return s.produced > b || s.eof; });
                    const uint64_t rows = s.produced > b ? s.rows[b % kRing] : 0;   // nothing after the end
                    if (rows < batch_size) batch_size = rows;                         // shortest prefix
                    share_ptrs[i] = s.ring[b % kRing];
                }
            }
            ms_wait += ms_since(t_wait);
            if (batch_size == 0) break;
            const clock::time_point t_combine = clock::now();
            double d;
            uint64_t idx;
            if (iris_combine_min(device, share_ptrs.data(), (uint32_t)parties, d_denominators + done * IRIS_ROTATIONS,
                                 batch_size, done, nullptr, &d, &idx))
                die("iris_combine_min");
            if (d < min_distance) {                                       // src/main.rs:614-617
                min_distance = d;
                min_index = idx;
            }
            done += batch_size;
            ms_combine += ms_since(t_combine);
            {
                std::lock_guard<std::mutex> lk(mu);
                consumed = b + 1;
            }
            cv.notify_all();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (Stream& s : streams) {
            shutdown(s.fd, SHUT_RDWR);                                    // unblocks a reader still inside read()
            s.reader.join();
            close(s.fd);
        }

        if (std::isinf(min_distance))
            fprintf(stderr, "Found closest entry at %" PRIu64 " out of %" PRIu64 " at distance inf.\n", min_index, done);
        else
            fprintf(stderr, "Found closest entry at %" PRIu64 " out of %" PRIu64 " at distance %.17g.\n", min_index, done,
                    min_distance);
        fprintf(stderr, "Timing: %.2f ms (connect + send %.2f, denominators %.2f, waiting for shares %.2f, combine %.2f)\n",
                ms_since(t_request), ms_connect, ms_denominators, ms_wait, ms_combine);
        printf("%" PRIu64 " %" PRIu64 " %.17g\n", min_index, done, min_distance);
        fflush(stdout);
    }
    if (qf) fclose(qf);
    for (Stream& s : streams)
        for (uint16_t* slot : s.ring) iris_host_free(slot);
    iris_device_free(device, d_denominators);
    iris_db_destroy(db);
    return 0;
}
