// Link-time replacement of ONE member function of the unmodified reference:
//   std::vector<LadderType> Board::GetLadderMap() const        (/root/reference/src/game/board.cc:1618-1688)
// — the ladder planes of every encoded position (Encoder::FillLadder, encoder.cc:248-266).  The front-end build
// (oracle/Makefile) weakens that symbol in its copy of board.o and links this file; the search itself is
// sayuri_b200/csrc/host_go/ladder.h, run on a 4 KB copy of exactly the arrays the reference's search reads (stones,
// neighbour counters, the four string arrays, the ko point).  Same result for every position (oracle/pass_alive_harness.cc
// `ladder`, tests/test_pass_alive.py).  Board::IsLadder stays the reference's for its other callers (patterns, GTP).
#include <cstdint>
#include <cstring>
#include <vector>

#include "game/board.h"

#include "../../host_go/ladder.h"

static_assert(sizeof(VertexType) == 1 && kNumVertices == sb_go::kLadderVertices && kMaxLadderNodes == sb_go::kLadderMaxNodes &&
                  (int)LadderType::kGoodForHunter == sb_go::kGoodForHunter && (int)LadderType::kNotLadder == sb_go::kNotLadder &&
                  (int)LadderType::kLadderTake == sb_go::kLadderTake && kNullVertex == 0,
              "coding of game/types.h");

std::vector<LadderType> Board::GetLadderMap() const {
    sb_go::LadderBoard b;
    std::memcpy(b.state, state_.data(), sizeof(b.state));
    std::memcpy(b.neighbours, neighbours_.data(), sizeof(b.neighbours));
    std::memcpy(b.next, strings_.next_.data(), sizeof(b.next));
    std::memcpy(b.parent, strings_.parent_.data(), sizeof(b.parent));
    std::memcpy(b.liberties, strings_.liberties_.data(), sizeof(b.liberties));
    std::memcpy(b.stones, strings_.stones_.data(), sizeof(b.stones));
    b.ko_move = ko_move_;
    b.board_size = board_size_;
    b.stride = letter_box_size_;
    for (int k = 0; k < 4; ++k) b.dir[k] = directions_[k];
    std::uint8_t map[kNumIntersections];
    sb_go::LadderMap(b, map);
    auto result = std::vector<LadderType>(num_intersections_);
    for (int i = 0; i < num_intersections_; ++i) result[i] = static_cast<LadderType>(map[i]);
    return result;
}
