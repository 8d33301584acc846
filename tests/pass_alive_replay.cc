// Replays tests/golden/pass_alive_cases.bin (the REFERENCE's answers, see tests/golden/make_pass_alive_golden.py) through
// sb_go::PassAliveArea and sb_go::ReachArea.  No reference code is needed to build or run this.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../sayuri_b200/csrc/host_go/pass_alive.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    long records = 0, answers = 0, mismatches = 0, marked = 0, reach_answers = 0;
    std::uint8_t header[2];
    while (std::fread(header, 1, 2, f) == 2) {
        const int n = header[0], cells = n * n, stride = n + 2;
        std::vector<std::uint8_t> stones(cells), state(stride * stride, sb_go::kInvalid), want(cells), got(cells);
        if ((int)std::fread(stones.data(), 1, cells, f) != cells) return 3;
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) state[(y + 1) * stride + x + 1] = stones[y * n + x];
        for (int color = 0; color < 2; ++color) {
            for (int flags = 0; flags < 4; ++flags) {
                if ((int)std::fread(want.data(), 1, cells, f) != cells) return 3;
                std::fill(got.begin(), got.end(), 0);
                sb_go::PassAliveArea(sb_go::BoardView{state.data(), n, stride}, color, flags & 1, flags & 2, got.data());
                ++answers;
                bool bad = false;
                for (int i = 0; i < cells; ++i) {
                    bad |= got[i] != want[i];
                    marked += want[i];
                }
                mismatches += bad;
            }
        }
        if ((int)std::fread(want.data(), 1, cells, f) != cells) return 3;
        std::vector<int> reach(cells, -1);
        sb_go::ReachArea(sb_go::BoardView{state.data(), n, stride}, reach.data());
        bool bad_reach = false;
        for (int i = 0; i < cells; ++i) bad_reach |= reach[i] != (int)want[i];
        mismatches += bad_reach;
        ++reach_answers;
        ++records;
    }
    std::fclose(f);
    std::printf("{\"records\": %ld, \"answers\": %ld, \"reach_answers\": %ld, \"marked_points\": %ld, \"mismatches\": %ld}\n", records, answers,
                reach_answers, marked, mismatches);
    return mismatches ? 1 : 0;
}
