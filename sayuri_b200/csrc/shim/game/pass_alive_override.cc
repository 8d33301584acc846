// Link-time replacement of TWO member functions of the unmodified reference:
//   void Board::ComputePassAliveArea(std::vector<bool>&, int, bool, bool) const   (/root/reference/src/game/board.cc:1720)
//   void Board::ComputeReachArea(std::vector<int>&) const                         (/root/reference/src/game/board.cc:1547)
// The front-end build (oracle/Makefile, targets sayuri_b200_frontend / sayuri_b200_det) weakens that symbol in its
// copy of board.o (objcopy --weaken-symbol) and links this file, so every caller — Board::ComputeSafeArea,
// Board::ComputeScoreArea and through them Encoder::FillArea and GameState::GetStrictSafeArea — gets the flat-array
// implementation of sayuri_b200/csrc/host_go/pass_alive.h.  A maintainer would paste the same three lines into
// board.cc.  Results are bit-identical (oracle/pass_alive_harness.cc, tests/test_pass_alive.py).
//
// The same position is asked three times per leaf (ComputeScoreArea and ComputeSafeArea back to back in
// Encoder::FillArea, encoder.cc:202-203, and GetStrictSafeArea when the node is expanded), always on the thread that
// expands the leaf: the last answer per colour is kept per thread and reused when the stones are byte-for-byte the
// same (the key is the whole state array, compared with memcmp — exact, no hashing).
#include <cstdint>
#include <cstring>
#include <vector>

#include "game/board.h"

#include "../../host_go/pass_alive.h"

static_assert(sizeof(VertexType) == 1 && kBlack == sb_go::kBlack && kWhite == sb_go::kWhite && kEmpty == sb_go::kEmpty &&
                  kInvalid == sb_go::kInvalid && kBoardSize <= sb_go::kMaxBoardSize,
              "vertex coding of game/types.h");

namespace {
struct LastAnswer {
    int board_size = 0;
    int flags = -1;
    std::uint8_t stones[kNumVertices];
    std::uint8_t out[kNumIntersections];
};
thread_local LastAnswer t_last[2];
}  // namespace

void Board::ComputePassAliveArea(std::vector<bool>& result, const int color, bool mark_vitals, bool mark_pass_dead) const {
    const auto* stones = reinterpret_cast<const std::uint8_t*>(state_.data());
    const int flags = (mark_vitals ? 1 : 0) | (mark_pass_dead ? 2 : 0);
    LastAnswer& last = t_last[color & 1];
    if (last.board_size != board_size_ || last.flags != flags || std::memcmp(last.stones, stones, (size_t)num_vertices_) != 0) {
        std::memset(last.out, 0, (size_t)num_intersections_);
        const sb_go::BoardView view{stones, board_size_, letter_box_size_};
        sb_go::PassAliveArea(view, color, mark_vitals, mark_pass_dead, last.out);
        std::memcpy(last.stones, stones, (size_t)num_vertices_);
        last.board_size = board_size_;
        last.flags = flags;
    }
    for (int i = 0; i < num_intersections_; ++i)
        if (last.out[i]) result[i] = true;
}

void Board::ComputeReachArea(std::vector<int>& result) const {
    if (result.size() != (size_t)num_intersections_) result.resize(num_intersections_);
    const sb_go::BoardView view{reinterpret_cast<const std::uint8_t*>(state_.data()), board_size_, letter_box_size_};
    sb_go::ReachArea(view, result.data());
}
