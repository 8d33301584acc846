// Link-time replacement of
//   void SelfPlayPipe::AssignDataWorker()   (/root/reference/src/selfplay/pipe.cc:181-233)
// — SURVEY.md §8(f) rank 4, "keep the data-writer off the hot threads".  The reference's writer thread polls its two
// queues in a `std::this_thread::yield()` loop for the whole run: one core (of the 16 that feed a B200) spent in
// sched_yield and in taking the two mutexes the game threads need (10 % of the samples and a third of the system time
// of a self-play profile, profiles/r01s2_selfplay_gprof.txt).  Finished games arrive a few times per second, so the
// replacement sleeps between polls.  What is written, when a chunk is due, the shuffle, the file formats (SaveChunk,
// SaveSgf, SaveNetQueries stay the reference's) and the shutdown handshake are unchanged:
//   * records are buffered until there are as many as parallel games, then written one by one, each time picking a
//     random survivor of a fresh shuffle; once the game threads are done (writing_worker_running_ == false)
//     everything left is written;
//   * the loop ends after a pass that saw the flag down and found both queues empty.
#include <algorithm>
#include <chrono>
#include <list>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "selfplay/pipe.h"
#include "utils/random.h"
#include "utils/threadpool.h"

void SelfPlayPipe::AssignDataWorker() {
    ThreadPool::Get("data-writer", 1);
    group_->AddTask([this]() -> void {
        constexpr float kValidationShare = 0.1f;
        constexpr auto kPollInterval = std::chrono::microseconds(500);
        const int chunk_after = engine_.GetParallelGames();
        std::vector<std::shared_ptr<DataSgfPair>> finished;
        std::list<GamesQueriesPair> query_lines;

        for (;;) {
            const bool games_running = writing_worker_running_.load(std::memory_order_relaxed);
            bool took_something = false;
            {
                std::lock_guard<std::mutex> hold(data_mutex_);
                took_something |= !data_sgf_buffer_.empty();
                for (auto& record : data_sgf_buffer_) finished.emplace_back(std::move(record));
                data_sgf_buffer_.clear();
            }
            {
                std::lock_guard<std::mutex> hold(log_mutex_);
                took_something |= !games_queries_buffer_.empty();
                query_lines.splice(query_lines.end(), games_queries_buffer_);
            }

            // the flag is read again, as the reference does: a run that ended while we drained flushes at once
            const size_t due = writing_worker_running_.load(std::memory_order_relaxed) ? (size_t)chunk_after : 1;
            while (finished.size() >= due) {
                std::shuffle(finished.begin(), finished.end(), Random<>::Get());
                const std::shared_ptr<DataSgfPair> record = finished.back();
                finished.pop_back();
                if (SaveChunk(num_saved_chunks_, kValidationShare, record->first)) ++num_saved_chunks_;
                SaveSgf(record->second);
            }
            for (auto& line : query_lines) SaveNetQueries(line.first, line.second);
            query_lines.clear();

            if (!games_running && !took_something) break;
            if (!took_something) std::this_thread::sleep_for(kPollInterval);
        }
    });
}
