// TEST INFRASTRUCTURE.  Links the UNMODIFIED reference objects (oracle/_ref/obj_v3) and compares, in one process,
//   Board::ComputePassAliveArea   /root/reference/src/game/board.cc:1720-1901   (private: reached with the usual
//                                                                                 `#define private public` test trick)
// with sb_go::PassAliveArea (sayuri_b200/csrc/host_go/pass_alive.h), and Board::ComputeReachArea (board.cc:1547-1579)
// with sb_go::ReachArea, on every position of seeded random games.
//   pass_alive_harness check <games> <seed>          all board sizes 2..19, both colours, the four flag combinations
//   pass_alive_harness time  <games> <seed> <size>   ns per call of each, (true, true) as all callers use
//   pass_alive_harness digest <games> <seed>         FNV-1a over Board::ComputeSafeArea and Board::ComputeScoreArea (public
//                                                    callers) on every position, each asked twice as the encoder does:
//                                                    the build with the link-time override (pass_alive_harness_fast,
//                                                    oracle/Makefile) must print the same digest as the plain one
//   pass_alive_harness encoder <games> <seed> [size] FNV-1a over the float bits of Encoder::GetPlanes (encoder.cc:31-50) for all
//                                                    eight symmetries on positions of random games played through
//                                                    GameState: every link-time override sits under that call, so
//                                                    plain and override builds must print the same digest; also
//                                                    prints the time per encoded position
//   pass_alive_harness ladder <games> <seed> [size]  Board::GetLadderMap (board.cc:1618-1688, whatever this binary links: the
//                                                    reference's in the plain build, the link-time override in the _fast
//                                                    build) against sb_go::LadderMap (host_go/ladder.h) on a copy of the
//                                                    board's string arrays, on every position of seeded random games;
//                                                    prints mismatches, an FNV digest of all maps and ns per map
//   pass_alive_harness dumpladder <games> <seed> <file>   ladder fixtures: board arrays + the map this binary's Board::GetLadderMap gives
//   pass_alive_harness encparts <games> <seed> <size>  us per call of each stage of Encoder::EncoderFeatures / GetPlanes (identity symmetry)
//   pass_alive_harness dump  <games> <seed> <file>   fixture file for tests/test_pass_alive.py (positions + the
//                                                    REFERENCE's answers), format in tests/test_pass_alive.py
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <array>
#include <functional>
#include <memory>
#include <sstream>
#include <iostream>

#define private public
#include "game/board.h"
#include "game/game_state.h"
#include "config.h"
#include "neural/encoder.h"
#undef private

#include "../sayuri_b200/csrc/host_go/pass_alive.h"
#include "../sayuri_b200/csrc/host_go/ladder.h"

namespace {

std::uint64_t SplitMix(std::uint64_t& s) {
    std::uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// One random game; `visit` sees the board after every move.  Mostly eye-respecting play (so that living groups and
// dead stones appear), sometimes not (so that eyes get filled and big captures happen).
template <typename F> void PlayGame(int size, std::uint64_t& rng, F&& visit) {
    Board board;
    board.Reset(size);
    int color = kBlack, passes = 0;
    const int max_moves = size * size * 3;
    const bool respect_eyes = SplitMix(rng) % 8 != 0;
    visit(board);
    for (int move = 0; move < max_moves && passes < 2; ++move) {
        std::vector<int> cand;
        for (int i = 0; i < board.GetEmptyCount(); ++i) {
            const int vtx = board.GetEmpty(i);
            if (!board.IsLegalMove(vtx, color)) continue;
            if (respect_eyes && board.IsRealEye(vtx, color)) continue;
            cand.push_back(vtx);
        }
        if (cand.empty() || SplitMix(rng) % 97 == 0) {
            board.PlayMoveAssumeLegal(kPass, color);
            ++passes;
        } else {
            board.PlayMoveAssumeLegal(cand[SplitMix(rng) % cand.size()], color);
            passes = 0;
        }
        color = !color;
        visit(board);
    }
}

sb_go::BoardView View(const Board& b) {
    return sb_go::BoardView{reinterpret_cast<const std::uint8_t*>(b.state_.data()), b.GetBoardSize(), b.GetLetterBoxSize()};
}

int SizeOfGame(int g) {   // small boards settle life and death quickly; big ones are what self-play runs
    static const int sizes[] = {2, 3, 4, 5, 5, 6, 7, 7, 8, 9, 9, 9, 10, 11, 12, 13, 13, 15, 17, 19, 19};
    return sizes[g % (int)(sizeof(sizes) / sizeof(sizes[0]))];
}

int Check(int games, std::uint64_t seed) {
    std::uint64_t rng = seed;
    long positions = 0, calls = 0, mismatches = 0, marked = 0, with_dead = 0;
    for (int g = 0; g < games; ++g) {
        PlayGame(SizeOfGame(g), rng, [&](const Board& b) {
            ++positions;
            const int n = b.GetNumIntersections();
            {
                std::vector<int> ref_reach(n, -1);
                int our_reach[kNumIntersections];
                b.ComputeReachArea(ref_reach);
                sb_go::ReachArea(View(b), our_reach);
                bool bad = false;
                for (int i = 0; i < n; ++i) bad |= ref_reach[i] != our_reach[i];
                if (bad && mismatches++ < 3) std::fprintf(stderr, "REACH MISMATCH size %d\n%s", b.GetBoardSize(), b.GetBoardString(kNullVertex, true).c_str());
                ++calls;
            }
            for (int color = 0; color < 2; ++color) {
                for (int flags = 0; flags < 4; ++flags) {
                    const bool vit = flags & 1, dead = flags & 2;
                    std::vector<bool> ref(n, false);
                    b.ComputePassAliveArea(ref, color, vit, dead);
                    std::uint8_t ours[kNumIntersections] = {0};
                    sb_go::PassAliveArea(View(b), color, vit, dead, ours);
                    ++calls;
                    bool bad = false;
                    for (int i = 0; i < n; ++i) {
                        bad |= (bool)ours[i] != (bool)ref[i];
                        marked += ref[i];
                    }
                    if (flags == 3) {
                        std::vector<bool> alive_only(n, false);
                        b.ComputePassAliveArea(alive_only, color, true, false);
                        with_dead += alive_only != ref;
                    }
                    if (bad && mismatches++ < 3) {
                        std::fprintf(stderr, "MISMATCH size %d color %d vitals %d dead %d\n%s", b.GetBoardSize(), color, (int)vit, (int)dead,
                                     b.GetBoardString(kNullVertex, true).c_str());
                        for (int y = b.GetBoardSize() - 1; y >= 0; --y) {
                            for (int x = 0; x < b.GetBoardSize(); ++x) std::fprintf(stderr, "%c", "._"[0] + 0 * x + (ref[y * b.GetBoardSize() + x] ? 'R' - '.' : 0));
                            std::fprintf(stderr, "   ");
                            for (int x = 0; x < b.GetBoardSize(); ++x) std::fprintf(stderr, "%c", ours[y * b.GetBoardSize() + x] ? 'O' : '.');
                            std::fprintf(stderr, "\n");
                        }
                    }
                }
            }
        });
    }
    std::printf("{\"games\": %d, \"positions\": %ld, \"calls\": %ld, \"marked_points\": %ld, \"calls_with_pass_dead_points\": %ld, \"mismatches\": %ld}\n",
                games, positions, calls, marked, with_dead, mismatches);
    return mismatches ? 1 : 0;
}

int Digest(int games, std::uint64_t seed) {
    std::uint64_t rng = seed, h = 0xcbf29ce484222325ull;
    auto mix = [&](std::uint64_t v) {
        h ^= v;
        h *= 0x100000001b3ull;
    };
    Board previous;
    previous.Reset(9);
    for (int g = 0; g < games; ++g) {
        PlayGame(SizeOfGame(g), rng, [&](const Board& b) {
            const int n = b.GetNumIntersections();
            std::vector<int> helper(n, kEmpty), score(n, kInvalid);
            std::vector<bool> safe(n, false);
            b.ComputeScoreArea(score, kArea, helper);
            b.ComputeSafeArea(safe, false);
            for (int i = 0; i < n; ++i) mix((std::uint64_t)score[i] * 2 + safe[i]);
            // interleave another board (the per-thread reuse must notice), then ask again
            std::vector<bool> other(previous.GetNumIntersections(), false);
            previous.ComputeSafeArea(other, false);
            for (size_t i = 0; i < other.size(); ++i) mix(other[i]);
            std::vector<bool> again(n, false);
            b.ComputeSafeArea(again, true);
            for (int i = 0; i < n; ++i) mix(again[i]);
            if (SplitMix(rng) % 3 == 0) previous = b;
        });
    }
    std::printf("%016llx\n", (unsigned long long)h);
    return 0;
}

int EncoderDigest(int games, std::uint64_t seed, int only_size) {
    std::uint64_t rng = seed, h = 0xcbf29ce484222325ull;
    long encoded = 0;
    double seconds = 0;
    for (int g = 0; g < games; ++g) {
        const int size = only_size ? only_size : SizeOfGame(g);
        GameState state;
        state.Reset(size, 7.5f - (g % 3), kArea);
        int passes = 0;
        for (int move = 0; move < size * size * 2 && passes < 2; ++move) {
            const int color = state.GetToMove();
            int vtx = kPass;
            for (int attempt = 0; attempt < 30; ++attempt) {
                const int x = SplitMix(rng) % size, y = SplitMix(rng) % size;
                const int v = state.GetVertex(x, y);
                if (state.IsLegalMove(v, color) && !state.board_.IsRealEye(v, color)) {
                    vtx = v;
                    break;
                }
            }
            state.PlayMove(vtx, color);
            passes = vtx == kPass ? passes + 1 : 0;
            if (move % 3) continue;
            for (int symm = 0; symm < 8; ++symm) {
                const auto t0 = std::chrono::steady_clock::now();
                const std::vector<float> planes = Encoder::Get().GetPlanes(state, symm, 5);
                seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                for (float f : planes) {
                    std::uint32_t bits;
                    std::memcpy(&bits, &f, 4);
                    h ^= bits;
                    h *= 0x100000001b3ull;
                }
                ++encoded;
            }
        }
    }
    std::printf("{\"digest\": \"%016llx\", \"encoded_positions\": %ld, \"us_per_position\": %.1f}\n", (unsigned long long)h, encoded,
                seconds * 1e6 / (encoded ? encoded : 1));
    return 0;
}

int Time(int games, std::uint64_t seed, int size) {
    std::uint64_t rng = seed;
    std::vector<Board> boards;
    for (int g = 0; g < games; ++g) PlayGame(size, rng, [&](const Board& b) { boards.push_back(b); });
    using clk = std::chrono::steady_clock;
    long sink = 0;
    const auto t0 = clk::now();
    for (const Board& b : boards) {
        for (int color = 0; color < 2; ++color) {
            std::vector<bool> ref(b.GetNumIntersections(), false);
            b.ComputePassAliveArea(ref, color, true, true);
            sink += ref[0];
        }
    }
    const auto t1 = clk::now();
    for (const Board& b : boards) {
        for (int color = 0; color < 2; ++color) {
            std::uint8_t ours[kNumIntersections] = {0};
            sb_go::PassAliveArea(View(b), color, true, true, ours);
            sink += ours[0];
        }
    }
    const auto t2 = clk::now();
    for (const Board& b : boards) {
        std::vector<int> reach(b.GetNumIntersections(), -1);
        b.ComputeReachArea(reach);
        sink += reach[0];
    }
    const auto t3 = clk::now();
    for (const Board& b : boards) {
        int reach[kNumIntersections];
        sb_go::ReachArea(View(b), reach);
        sink += reach[0];
    }
    const auto t4 = clk::now();
    const double n = 2.0 * boards.size();
    std::printf("{\"board_size\": %d, \"positions\": %zu, \"reference_ns_per_call\": %.0f, \"ours_ns_per_call\": %.0f, "
                "\"reach_reference_ns_per_call\": %.0f, \"reach_ours_ns_per_call\": %.0f, \"sink\": %ld}\n", size,
                boards.size(), std::chrono::duration<double, std::nano>(t1 - t0).count() / n,
                std::chrono::duration<double, std::nano>(t2 - t1).count() / n,
                std::chrono::duration<double, std::nano>(t3 - t2).count() / boards.size(),
                std::chrono::duration<double, std::nano>(t4 - t3).count() / boards.size(), sink);
    return 0;
}

int Dump(int games, std::uint64_t seed, const char* path) {
    std::uint64_t rng = seed;
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return 2;
    long records = 0;
    for (int g = 0; g < games; ++g) {
        int k = 0;
        PlayGame(SizeOfGame(g), rng, [&](const Board& b) {
            // every 5th position plus the late ones (where life and death is settled)
            if (k++ % 5 && SplitMix(rng) % 4) return;
            const int n = b.GetBoardSize(), cells = n * n;
            std::uint8_t header[2] = {(std::uint8_t)n, 0};
            std::fwrite(header, 1, 2, f);
            std::vector<std::uint8_t> stones(cells);
            for (int i = 0; i < cells; ++i) stones[i] = (std::uint8_t)b.GetState(b.IndexToVertex(i));
            std::fwrite(stones.data(), 1, cells, f);
            for (int color = 0; color < 2; ++color) {
                for (int flags = 0; flags < 4; ++flags) {
                    std::vector<bool> ref(cells, false);
                    b.ComputePassAliveArea(ref, color, flags & 1, flags & 2);
                    std::vector<std::uint8_t> bytes(cells);
                    for (int i = 0; i < cells; ++i) bytes[i] = ref[i];
                    std::fwrite(bytes.data(), 1, cells, f);
                }
            }
            std::vector<int> reach(cells, -1);
            b.ComputeReachArea(reach);
            std::vector<std::uint8_t> reach_bytes(cells);
            for (int i = 0; i < cells; ++i) reach_bytes[i] = (std::uint8_t)reach[i];
            std::fwrite(reach_bytes.data(), 1, cells, f);
            ++records;
        });
    }
    std::fclose(f);
    std::printf("%ld records\n", records);
    return 0;
}

}  // namespace

sb_go::LadderBoard LadderView(const Board& b) {
    sb_go::LadderBoard lb;
    std::memcpy(lb.state, b.state_.data(), sizeof(lb.state));
    std::memcpy(lb.neighbours, b.neighbours_.data(), sizeof(lb.neighbours));
    std::memcpy(lb.next, b.strings_.next_.data(), sizeof(lb.next));
    std::memcpy(lb.parent, b.strings_.parent_.data(), sizeof(lb.parent));
    std::memcpy(lb.liberties, b.strings_.liberties_.data(), sizeof(lb.liberties));
    std::memcpy(lb.stones, b.strings_.stones_.data(), sizeof(lb.stones));
    lb.ko_move = b.ko_move_;
    lb.board_size = b.board_size_;
    lb.stride = b.letter_box_size_;
    for (int k = 0; k < 4; ++k) lb.dir[k] = b.directions_[k];
    return lb;
}

// Ladders need fights: random play with a bias towards ataris (a move next to a string with two liberties)
template <typename F> void PlayLadderGame(int size, std::uint64_t& rng, F&& visit) {
    Board board;
    board.Reset(size);
    int color = kBlack, passes = 0;
    const int max_moves = size * size * 2;
    visit(board);
    for (int move = 0; move < max_moves && passes < 2; ++move) {
        std::vector<int> cand, sharp;
        for (int i = 0; i < board.GetEmptyCount(); ++i) {
            const int vtx = board.GetEmpty(i);
            if (!board.IsLegalMove(vtx, color) || board.IsRealEye(vtx, color)) continue;
            cand.push_back(vtx);
            for (int k = 0; k < 4; ++k) {
                const int a = vtx + board.directions_[k];
                if (board.state_[a] == !color && board.strings_.GetLiberty(board.strings_.GetParent(a)) <= 2) {
                    sharp.push_back(vtx);
                    break;
                }
            }
        }
        if (cand.empty() || SplitMix(rng) % 131 == 0) {
            board.PlayMoveAssumeLegal(kPass, color);
            ++passes;
        } else {
            const auto& from = (!sharp.empty() && SplitMix(rng) % 3 != 0) ? sharp : cand;
            board.PlayMoveAssumeLegal(from[SplitMix(rng) % from.size()], color);
            passes = 0;
        }
        color = !color;
        visit(board);
    }
}

int Ladder(int games, std::uint64_t seed, int only_size, const char* dump_path) {
    std::uint64_t rng = seed;
    long positions = 0, mismatches = 0, marked = 0, with_ladder = 0;
    std::uint64_t digest = 1469598103934665603ull;
    double t_member = 0, t_header = 0;
    FILE* f = dump_path ? std::fopen(dump_path, "wb") : nullptr;
    static const int sizes[] = {5, 7, 9, 9, 11, 13, 13, 15, 17, 19, 19, 19};
    for (int g = 0; g < games; ++g) {
        const int size = only_size ? only_size : sizes[g % 12];
        PlayLadderGame(size, rng, [&](const Board& b) {
            ++positions;
            const int n = b.GetNumIntersections();
            const auto t0 = std::chrono::steady_clock::now();
            const std::vector<LadderType> member = b.GetLadderMap();
            const auto t1 = std::chrono::steady_clock::now();
            const sb_go::LadderBoard lb = LadderView(b);
            std::uint8_t ours[kNumIntersections];
            sb_go::LadderMap(lb, ours);
            const auto t2 = std::chrono::steady_clock::now();
            t_member += std::chrono::duration<double>(t1 - t0).count();
            t_header += std::chrono::duration<double>(t2 - t1).count();
            bool bad = false, any = false;
            for (int i = 0; i < n; ++i) {
                bad |= (int)member[i] != (int)ours[i];
                any |= member[i] != LadderType::kNotLadder;
                marked += member[i] != LadderType::kNotLadder;
                digest = (digest ^ (std::uint64_t)member[i]) * 1099511628211ull;
            }
            with_ladder += any;
            if (bad && mismatches++ < 3) {
                std::fprintf(stderr, "LADDER MISMATCH size %d\n%s", b.GetBoardSize(), b.GetBoardString(kNullVertex, true).c_str());
                for (int y = b.GetBoardSize() - 1; y >= 0; --y) {
                    for (int x = 0; x < b.GetBoardSize(); ++x) std::fprintf(stderr, "%d", (int)member[y * b.GetBoardSize() + x]);
                    std::fprintf(stderr, "   ");
                    for (int x = 0; x < b.GetBoardSize(); ++x) std::fprintf(stderr, "%d", (int)ours[y * b.GetBoardSize() + x]);
                    std::fprintf(stderr, "\n");
                }
            }
            if (f && any && (positions % 7 == 0)) {   // a sample of the positions that do contain a ladder
                std::fwrite(&lb, sizeof(lb), 1, f);
                std::uint8_t exp[kNumIntersections] = {0};
                for (int i = 0; i < n; ++i) exp[i] = (std::uint8_t)member[i];
                std::fwrite(exp, 1, kNumIntersections, f);
            }
        });
    }
    if (f) std::fclose(f);
    std::printf("{\"games\": %d, \"positions\": %ld, \"positions_with_a_ladder\": %ld, \"marked_points\": %ld, \"mismatches\": %ld, "
                "\"digest\": \"%016llx\", \"ns_per_map_member\": %.0f, \"ns_per_map_header\": %.0f}\n",
                games, positions, with_ladder, marked, mismatches, (unsigned long long)digest, 1e9 * t_member / positions, 1e9 * t_header / positions);
    return mismatches ? 1 : 0;
}

int EncoderParts(int games, std::uint64_t seed, int size) {
    std::uint64_t rng = seed;
    double t[8] = {0};
    long n = 0;
    const Encoder& enc = Encoder::Get();
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    for (int g = 0; g < games; ++g) {
        GameState state;
        state.Reset(size, 7.5f, kArea);
        int passes = 0;
        for (int move = 0; move < size * size * 2 && passes < 2; ++move) {
            const int color = state.GetToMove();
            int vtx = kPass;
            for (int attempt = 0; attempt < 30; ++attempt) {
                const int v = state.GetVertex(SplitMix(rng) % size, SplitMix(rng) % size);
                if (state.IsLegalMove(v, color) && !state.board_.IsRealEye(v, color)) {
                    vtx = v;
                    break;
                }
            }
            state.PlayMove(vtx, color);
            passes = vtx == kPass ? passes + 1 : 0;
            const int ns = size * size;
            std::vector<float> planes(43 * ns, 0.f);
            auto it = planes.begin();
            auto board = state.GetPastBoard(0);
            auto a = now(); enc.EncoderHistoryMove(state, it, 5); t[0] += since(a);
            a = now(); enc.FillKoMove(board.get(), it + 24 * ns); t[1] += since(a);
            a = now(); enc.FillArea(board.get(), color, kArea, it + 25 * ns, 5); t[2] += since(a);
            a = now(); enc.FillLiberties(board.get(), it + 29 * ns); t[3] += since(a);
            a = now(); enc.FillLadder(board.get(), it + 33 * ns); t[4] += since(a);
            a = now(); enc.FillMisc(board.get(), color, kArea, 0.f, 7.5f, it + 37 * ns, 5); t[5] += since(a);
            a = now(); { const std::vector<float> p2 = enc.GetPlanes(state, 0, 5); t[6] += since(a); }
            a = now(); { const InputData in = enc.GetInputs(state, 0, 5); t[7] += since(a); }
            ++n;
        }
    }
    const char* names[8] = {"EncoderHistoryMove", "FillKoMove", "FillArea", "FillLiberties", "FillLadder", "FillMisc", "GetPlanes(total)", "GetInputs(total)"};
    std::printf("{\"positions\": %ld", n);
    for (int i = 0; i < 8; ++i) std::printf(", \"%s_us\": %.2f", names[i], 1e6 * t[i] / n);
    std::printf("}\n");
    return 0;
}

int main(int argc, char** argv) {
    static char a0[] = "pass_alive_harness", a1[] = "--quiet";
    char* args[] = {a0, a1};
    ArgsParser(2, args);   // Zobrist tables etc. (config.cc:336-381)
    if (argc >= 4 && !std::strcmp(argv[1], "check")) return Check(std::atoi(argv[2]), std::strtoull(argv[3], nullptr, 10));
    if (argc >= 4 && !std::strcmp(argv[1], "encoder")) return EncoderDigest(std::atoi(argv[2]), std::strtoull(argv[3], nullptr, 10), argc >= 5 ? std::atoi(argv[4]) : 0);
    if (argc >= 4 && !std::strcmp(argv[1], "digest")) return Digest(std::atoi(argv[2]), std::strtoull(argv[3], nullptr, 10));
    if (argc >= 5 && !std::strcmp(argv[1], "time")) return Time(std::atoi(argv[2]), std::strtoull(argv[3], nullptr, 10), std::atoi(argv[4]));
    if (argc >= 5 && !std::strcmp(argv[1], "encparts")) return EncoderParts(std::atoi(argv[2]), std::strtoull(argv[3], nullptr, 10), std::atoi(argv[4]));
    if (argc >= 4 && !std::strcmp(argv[1], "ladder")) return Ladder(std::atoi(argv[2]), std::strtoull(argv[3], nullptr, 10), argc >= 5 ? std::atoi(argv[4]) : 0, nullptr);
    if (argc >= 5 && !std::strcmp(argv[1], "dumpladder")) return Ladder(std::atoi(argv[2]), std::strtoull(argv[3], nullptr, 10), 0, argv[4]);
    if (argc >= 5 && !std::strcmp(argv[1], "dump")) return Dump(std::atoi(argv[2]), std::strtoull(argv[3], nullptr, 10), argv[4]);
    std::fprintf(stderr, "usage: pass_alive_harness check <games> <seed> | time <games> <seed> <size> | dump <games> <seed> <file>\n");
    return 2;
}
