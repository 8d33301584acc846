// TEST INFRASTRUCTURE.  One source, two builds (oracle/Makefile): with -I $REF/src only it instantiates the
// REFERENCE's HashKeyCache (/root/reference/src/utils/cache.h), with -I sayuri_b200/csrc/shim first it instantiates
// ours.  Neither build copies reference code.
//   cache_harness parity <n_ops> <seed> <capacity>   single-threaded random op sequence -> one FNV-1a digest of every
//                                                    observable (hit/miss, the value returned); the two builds must
//                                                    print the same digest
//   cache_harness bench <threads> <seconds> <capacity> <key_space> <probes>
//                                                    every thread loops: `probes` lookups of related keys (as
//                                                    Network::ProbeCache does in the opening, network.cc:197-235),
//                                                    insert on a miss (network.cc:286); prints evaluations/s
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "utils/cache.h"

namespace {

// Same size class as Network::Result (two 361-float arrays + scalars, network_basic.h:36-63).
struct Value {
    float body[2 * 361 + 14];
    std::uint64_t tag;
};

inline std::uint64_t SplitMix(std::uint64_t& s) {
    std::uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

inline void Fnv(std::uint64_t& h, std::uint64_t v) {
    for (int i = 0; i < 8; ++i) {
        h ^= (v >> (8 * i)) & 0xff;
        h *= 0x100000001b3ull;
    }
}

int Parity(long n_ops, std::uint64_t seed, size_t capacity) {
    HashKeyCache<Value> cache(capacity);
    std::uint64_t rng = seed, digest = 0xcbf29ce484222325ull;
    Fnv(digest, cache.GetEntrySize());
    const std::uint64_t key_space = capacity * 3 + 7;   // enough pressure for evictions, enough reuse for hits
    Value v, got;
    for (long i = 0; i < n_ops; ++i) {
        const std::uint64_t r = SplitMix(rng);
        // keys that collide on the cluster index but differ above it, and plain ones
        std::uint64_t key = SplitMix(rng) % key_space;
        if ((r & 7) == 0) key += (capacity / 8 ? capacity / 8 : 1) * (1 + (r >> 8) % 9);
        const unsigned op = (r >> 3) % 100;
        if (op < 55) {
            std::memset(&got, 0, sizeof(got));
            const bool hit = cache.LookupItem(key, got);
            Fnv(digest, hit ? 1 : 0);
            if (hit) {
                Fnv(digest, got.tag);
                std::uint32_t bits;
                std::memcpy(&bits, &got.body[i % 736], 4);
                Fnv(digest, bits);
            }
        } else if (op < 99) {
            v.tag = r;
            for (int k = 0; k < 736; ++k) v.body[k] = (float)((r >> (k & 31)) & 1023) * 0.25f;
            cache.Insert(key, v);
        } else if ((r >> 20) % 50 == 0) {
            cache.Clear();
        } else if ((r >> 20) % 50 == 1) {
            cache.SetCapacity(capacity + (r >> 40) % 64);   // the reference grows in place: later clusters start empty
        }
    }
    HashKeyCache<Value> moved(std::move(cache));         // empty, same capacity, continues the generation count
    Fnv(digest, moved.LookupItem(1, got) ? 1 : 0);
    v.tag = 99;
    moved.Insert(1, v);
    Fnv(digest, moved.LookupItem(1, got) ? got.tag : 0);
    std::printf("%016llx\n", (unsigned long long)digest);
    return 0;
}

int Bench(int threads, double seconds, size_t capacity, std::uint64_t key_space, int probes) {
    HashKeyCache<Value> cache(capacity);
    std::atomic<bool> stop{false};
    std::atomic<long> evals{0}, hits{0}, torn{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t] {
            std::uint64_t rng = 0x1234 + 7919ull * t;
            Value v, got;
            std::memset(&v, 0, sizeof(v));
            long n = 0, h = 0, bad = 0;
            while (!stop.load(std::memory_order_relaxed)) {
                const std::uint64_t key = SplitMix(rng) % key_space;
                bool hit = false;
                for (int p = probes - 1; p >= 0 && !hit; --p) hit = cache.LookupItem(key ^ ((std::uint64_t)p << 56), got);
                if (!hit) {
                    v.tag = key;
                    for (int k = 0; k < 736; k += 61) v.body[k] = (float)(key & 0xffff);
                    cache.Insert(key, v);
                } else {
                    ++h;
                    // a hit must return a whole value written for this key (only probe 0 is ever inserted)
                    if (got.tag != key) ++bad;
                    for (int k = 0; k < 736; k += 61) bad += got.body[k] != (float)(key & 0xffff);
                }
                ++n;
            }
            evals += n;
            hits += h;
            torn += bad;
        });
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
    stop = true;
    for (auto& th : pool) th.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("{\"threads\": %d, \"probes\": %d, \"evals_per_s\": %.0f, \"hit_rate\": %.3f, \"torn\": %ld}\n", threads, probes,
                evals.load() / dt, evals.load() ? (double)hits.load() / evals.load() : 0.0, torn.load());
    return torn.load() ? 1 : 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc >= 5 && !std::strcmp(argv[1], "parity")) return Parity(std::atol(argv[2]), std::strtoull(argv[3], nullptr, 10), std::strtoull(argv[4], nullptr, 10));
    if (argc >= 7 && !std::strcmp(argv[1], "bench"))
        return Bench(std::atoi(argv[2]), std::atof(argv[3]), std::strtoull(argv[4], nullptr, 10), std::strtoull(argv[5], nullptr, 10), std::atoi(argv[6]));
    std::fprintf(stderr, "usage: cache_harness parity <n_ops> <seed> <capacity> | bench <threads> <seconds> <capacity> <key_space> <probes>\n");
    return 2;
}
