// Drop-in replacement of the reference's NN result cache (/root/reference/src/utils/cache.h:10-129, used as
// Network::Cache = HashKeyCache<Result>, src/neural/network.h:23,67) — SURVEY.md §8(f) rank 3.  Shadows the reference
// header by include order (-I sayuri_b200/csrc/shim before -I $REF/src), exactly like neural/cuda/cuda_forward_pipe.h.
//
// Why: the reference guards the whole table with ONE spin lock and, inside it, allocates and copies a ~3 KB result on
// every insert and copies one on every hit.  Every leaf evaluation takes that lock 2..9 times (ProbeCache probes up to
// eight symmetry hashes in the opening, network.cc:197-235, then Insert, :286): at the rates this engine delivers
// (>1 M evals/s on 8 GPUs) the lock alone is the ceiling (measured: oracle/cache_harness.cc, profiles/r01s2_cache_bench.md).
//
// Same observable behaviour, different synchronisation:
//  * the table is the same array of 8-entry clusters, cluster = key % n_clusters; a lookup returns the first entry of
//    the cluster (in slot order) whose key matches and that has ever been written; an insert replaces the entry of the
//    cluster with the smallest generation stamp (first such slot on ties) — never checks for an existing key;
//  * generation stamps come from one atomic counter (fetch_add), so within a cluster they order inserts exactly as
//    the reference's locked counter does: a single-threaded sequence of operations gives bit-identical hits, misses
//    and evictions (tests/test_cache.py runs both headers on the same sequences);
//  * one lock per STRIPE of clusters (1024 cache-line-padded locks) instead of one per table; the allocation and the
//    copy of an inserted value, and the release of the evicted one, happen outside any lock; the copy-out of a hit
//    happens under the stripe lock only;
//  * SetCapacity / Clear take every stripe (the reference takes its single lock).
//
// -DSAYURI_B200_REF_CACHE turns this header into a pass-through to the reference's own (A/B builds, oracle/Makefile).
#pragma once

#ifdef SAYURI_B200_REF_CACHE
#include_next "utils/cache.h"
#else

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#endif

template <typename V> class HashKeyCache {
public:
    HashKeyCache() = default;
    explicit HashKeyCache(size_t capacity) { SetCapacity(capacity); }

    // Same meaning as the reference's move constructor: an EMPTY table of the same capacity that continues the
    // generation count (cache.h:22-26).
    HashKeyCache(HashKeyCache&& other) {
        SetCapacity(other.capacity_);
        Clear();
        next_stamp_.store(other.next_stamp_.load(std::memory_order_relaxed), std::memory_order_relaxed);
    }

    void SetCapacity(size_t size) {
        const size_t clusters = (size + kSlotsPerCluster - 1) / kSlotsPerCluster;
        AllStripes hold(*this);
        n_clusters_ = clusters;
        capacity_ = clusters * kSlotsPerCluster;
        slots_.resize(capacity_);
        slots_.shrink_to_fit();
    }

    void Insert(std::uint64_t key, const V& value) {
        if (n_clusters_ == 0) return;
        std::unique_ptr<V> fresh = std::make_unique<V>(value);   // allocate + copy before taking the stripe
        const size_t cluster = key % n_clusters_;
        Slot* const first = slots_.data() + cluster * kSlotsPerCluster;
        {
            StripeGuard hold(StripeOf(cluster));
            Slot* victim = first;
            for (size_t i = 1; i < kSlotsPerCluster; ++i)
                if (first[i].stamp < victim->stamp) victim = first + i;
            victim->key = key;
            victim->stamp = next_stamp_.fetch_add(1, std::memory_order_relaxed) + 1;
            victim->value.swap(fresh);
        }
        // `fresh` now owns the evicted value: freed here, outside the lock
    }

    bool LookupItem(std::uint64_t key, V& out) {
        if (n_clusters_ == 0) return false;
        const size_t cluster = key % n_clusters_;
        const Slot* const first = slots_.data() + cluster * kSlotsPerCluster;
        StripeGuard hold(StripeOf(cluster));
        for (size_t i = 0; i < kSlotsPerCluster; ++i) {
            if (first[i].stamp != 0 && first[i].key == key) {
                out = *first[i].value;
                return true;
            }
        }
        return false;
    }

    void Clear() {
        AllStripes hold(*this);
        next_stamp_.store(0, std::memory_order_relaxed);
        for (Slot& s : slots_) {
            s.stamp = 0;
            s.value.reset();
        }
    }

    size_t GetEntrySize() const { return sizeof(Slot) + sizeof(V); }

private:
    static constexpr size_t kSlotsPerCluster = 8;
    static constexpr size_t kStripes = 1024;   // power of two

    struct Slot {
        std::uint64_t key = 0;
        std::uint64_t stamp = 0;   // 0 = never written
        std::unique_ptr<V> value;
    };

    // Test-and-test-and-set lock on its own cache line; yields when it spins long (search threads outnumber cores).
    struct alignas(64) Stripe {
        std::atomic<std::uint32_t> held{0};
        void lock() {
            for (unsigned spins = 0;;) {
                if (!held.load(std::memory_order_relaxed) && !held.exchange(1, std::memory_order_acquire)) return;
                if (++spins & 63) {
#if defined(__x86_64__) || defined(_M_X64)
                    _mm_pause();
#endif
                } else {
                    std::this_thread::yield();
                }
            }
        }
        void unlock() { held.store(0, std::memory_order_release); }
    };
    struct StripeGuard {
        explicit StripeGuard(Stripe& s) : s_(s) { s_.lock(); }
        ~StripeGuard() { s_.unlock(); }
        StripeGuard(const StripeGuard&) = delete;
        StripeGuard& operator=(const StripeGuard&) = delete;
        Stripe& s_;
    };
    struct AllStripes {
        explicit AllStripes(HashKeyCache& c) : c_(c) {
            for (size_t i = 0; i < kStripes; ++i) c_.stripes_[i].lock();
        }
        ~AllStripes() {
            for (size_t i = kStripes; i-- > 0;) c_.stripes_[i].unlock();
        }
        HashKeyCache& c_;
    };

    Stripe& StripeOf(size_t cluster) { return stripes_[cluster & (kStripes - 1)]; }

    std::unique_ptr<Stripe[]> stripes_{new Stripe[kStripes]};
    std::vector<Slot> slots_;
    size_t capacity_ = 0;
    size_t n_clusters_ = 0;
    std::atomic<std::uint64_t> next_stamp_{0};
};

#endif  // SAYURI_B200_REF_CACHE
