// Replays tests/golden/ladder_cases.bin — board arrays of positions from seeded fight-heavy random games (stones, neighbour
// counters, the four string arrays, ko point: struct sb_go::LadderBoard) followed by the REFERENCE's Board::GetLadderMap
// answer (361 bytes, LadderType per intersection; written by `oracle/_ref/pass_alive_harness dumpladder` with the unmodified
// reference linked) — through sb_go::LadderMap.  No reference code is needed to build or run this.
#include <cstdint>
#include <cstdio>

#include "../sayuri_b200/csrc/host_go/ladder.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    long records = 0, mismatches = 0, marked = 0, sizes[32] = {0};
    sb_go::LadderBoard b;
    std::uint8_t want[361], got[361];
    while (std::fread(&b, sizeof(b), 1, f) == 1) {
        if (std::fread(want, 1, 361, f) != 361) return 3;
        if (b.board_size < 2 || b.board_size > 19 || b.stride != b.board_size + 2) return 4;
        sb_go::LadderMap(b, got);
        bool bad = false;
        for (int i = 0; i < b.board_size * b.board_size; ++i) {
            bad |= got[i] != want[i];
            marked += want[i] != sb_go::kNotLadder;
        }
        mismatches += bad;
        sizes[b.board_size]++;
        ++records;
    }
    std::fclose(f);
    std::printf("{\"records\": %ld, \"marked_points\": %ld, \"records_19x19\": %ld, \"mismatches\": %ld}\n", records, marked, sizes[19], mismatches);
    return mismatches ? 1 : 0;
}
