// Link-time replacements of three plane-filling members of the reference's Encoder:
//   void Encoder::SymmetryPlanes(const GameState&, std::vector<float>&, int, int) const   (/root/reference/src/neural/encoder.cc:80-100)
//   void Encoder::FillColorStones(const Board*, iterator black, iterator white) const     (encoder.cc:102-117, 8 x per position)
//   void Encoder::FillMove(const Board*, iterator) const                                  (encoder.cc:119-134, 8 x per position)
// — SURVEY.md §8 a6, the symmetry GATHER buf[i] = plane[T(i)] applied to all 43 input planes of every encoded position.
// The reference asks the Symmetry singleton for T(i) once per plane element (43 x 361 out-of-line Symmetry::Get()
// calls per leaf); the map depends only on (board size, symmetry), so it is read once per call and reused for every
// plane, and the identity symmetry (one leaf in eight under the random ensemble, all of them under kDirect) returns at
// once.  Same floats at the same places: oracle/pass_alive_harness.cc `encoder` digests, plain build vs this build.
#include <cstring>
#include <vector>

#include "game/board.h"
#include "game/game_state.h"
#include "game/symmetry.h"
#include "neural/encoder.h"

void Encoder::SymmetryPlanes(const GameState& state, std::vector<float>& planes, const int symmetry, const int weights_version) const {
    if (symmetry == Symmetry::kIdentitySymmetry) return;   // T(i) = i (symmetry.cc:97-123)
    const int board_size = state.GetBoardSize();
    const int cells = state.GetNumIntersections();
    const int channels = GetInputChannels(weights_version);
    const Symmetry& table = Symmetry::Get();
    int source[kNumIntersections];
    for (int i = 0; i < cells; ++i) source[i] = table.TransformIndex(board_size, symmetry, i);
    float turned[kNumIntersections];
    float* plane = planes.data();
    for (int c = 0; c < channels; ++c, plane += cells) {
        for (int i = 0; i < cells; ++i) turned[i] = plane[source[i]];
        std::memcpy(plane, turned, sizeof(float) * (size_t)cells);
    }
}

// Stones of the two colours as {0,1} planes: row-wise over the letter box instead of an index -> vertex division per
// point.  Only ones are written (the planes arrive zeroed and the reference writes nothing else either).
void Encoder::FillColorStones(const Board* board, std::vector<float>::iterator black_it, std::vector<float>::iterator white_it) const {
    const int n = board->GetBoardSize();
    float* const black = &*black_it;
    float* const white = &*white_it;
    for (int y = 0; y < n; ++y) {
        int vtx = board->GetVertex(0, y);
        for (int x = 0; x < n; ++x, ++vtx) {
            const int s = board->GetState(vtx);
            if (s == kBlack) {
                black[y * n + x] = 1.0f;
            } else if (s == kWhite) {
                white[y * n + x] = 1.0f;
            }
        }
    }
}

// One-hot plane of the last move: the reference scans all points for the one whose vertex equals the move.
void Encoder::FillMove(const Board* board, std::vector<float>::iterator move_it) const {
    const int last = board->GetLastMove();
    if (last == kNullVertex || last == kPass || last == kResign) return;
    const int n = board->GetBoardSize();
    const int x = board->GetX(last), y = board->GetY(last);
    if (x < 0 || x >= n || y < 0 || y >= n) return;   // not a point of the board: the scan would find nothing
    if (board->GetVertex(x, y) != last) return;
    move_it[board->GetIndex(x, y)] = 1.0f;
}
