// Pass-alive / pass-dead area of one colour on a Go board — the set that the reference computes in
// Board::ComputePassAliveArea (/root/reference/src/game/board.cc:1720-1901, helpers :1903-2175) — SURVEY.md §8(f)
// rank 2 ("Encoder on the critical path: Board::ComputeSafeArea").
//
// Why it is here: the reference evaluates that function 6.4 times per NN evaluation (twice per colour from
// Encoder::FillArea, encoder.cc:186-203 via ComputeScoreArea + ComputeSafeArea, and again from
// GameState::GetStrictSafeArea, game_state.cc:834), each time with ~40 heap allocations, std::function flood fills,
// std::vector<bool> scans of the whole letter box per group and a std::set per string; in a gprof run of the
// reference's self-play loop it is ~55 % of the host time outside the network (profiles/r01s2_selfplay_gprof.txt),
// and host time is what bounds games/hour once the network runs on a B200.
//
// This is a restatement of the same DEFINITION with flat stack arrays (no allocation, one labelling pass per
// component set): same result bit for bit, including the reference's quirks —
//   * Benson's algorithm with "vital region" = region of non-`color` points in which every EMPTY point touches a
//     `color` stone, healthy for a string when every empty point of it touches THAT string; a string needs two
//     (board.cc:1742-1840, 1903-1955; suicide is never allowed, :1738).  The fixed point does not depend on the order
//     in which strings are removed, so no order is copied;
//   * vital regions count as `color` for the pass-dead scan and removed strings as empty (:1862-1873, :1809);
//   * pass-dead = fewer than two potential eyes, two adjacent ones counting as one (:1957-2051); a diagonal that lies
//     in an "inner region" of the scanned region counts for the eye's owner (:2000-2006);
//   * inner regions (:2053-2107): components of the rest of the board that do not touch the edge — AND every
//     component that follows an edge-touching one in scan order, because the reference's erase-while-iterating loop
//     skips it (:2079-2097).  Kept, since it changes results.
// The parity harness (oracle/pass_alive_harness.cc) compares this function with the reference's on every position of
// seeded random games, all board sizes, both colours, all flag combinations; tests/test_pass_alive.py replays committed
// fixtures without the reference.
//
// Header-only, no dependency on the reference: `state` uses its vertex coding (0 black, 1 white, 2 empty, 3 off-board,
// game/types.h:36-50) on a letter-boxed board (vertex = (y + 1) * stride + x + 1, board.h:479-483).
#pragma once

#include <cstdint>
#include <cstring>

namespace sb_go {

constexpr int kBlack = 0, kWhite = 1, kEmpty = 2, kInvalid = 3;
constexpr int kMaxBoardSize = 25;                                   // kMaxGTPBoardSize, game/types.h:22
constexpr int kMaxVertices = (kMaxBoardSize + 2) * (kMaxBoardSize + 2);

struct BoardView {
    const std::uint8_t* state;   // [stride * stride]
    int board_size;
    int stride;                  // letter box width = board_size + 2
};

namespace detail {

// Connected components (4-neighbourhood) of the on-board vertices with feat[v] == target, numbered from 1 in scan
// order (y, then x) of their first vertex — the order of Board::ClassifyGroups' head list (board.cc:2109-2163).
struct Components {
    std::int16_t label[kMaxVertices];      // 0 = not a member of any component
    std::int16_t member[kMaxVertices];     // vertices grouped by component
    std::int16_t begin[kMaxVertices + 2];  // component c owns member[begin[c] .. begin[c + 1])
    int count = 0;

    void Build(const std::uint8_t* feat, int target, int n, int stride) {
        std::memset(label, 0, sizeof(label[0]) * stride * stride);
        count = 0;
        int filled = 0;
        begin[1] = 0;
        const int d[4] = {-stride, -1, +1, +stride};
        for (int y = 0; y < n; ++y) {
            for (int x = 0; x < n; ++x) {
                const int v0 = (y + 1) * stride + x + 1;
                if (label[v0] || feat[v0] != target) continue;
                const int c = ++count;
                int head = filled;                 // member[] doubles as the flood-fill queue
                label[v0] = (std::int16_t)c;
                member[filled++] = (std::int16_t)v0;
                while (head < filled) {
                    const int v = member[head++];
                    for (int k = 0; k < 4; ++k) {
                        const int a = v + d[k];
                        if (!label[a] && feat[a] == target) {   // off-board vertices carry kInvalid: never the target
                            label[a] = (std::int16_t)c;
                            member[filled++] = (std::int16_t)a;
                        }
                    }
                }
                begin[c + 1] = (std::int16_t)filled;
            }
        }
    }
};

}  // namespace detail

// out[y * board_size + x] is set to 1 for every point the reference would set to true; other entries are untouched
// (the reference ORs into its result the same way, board.cc:1846,1861,1885).
inline void PassAliveArea(const BoardView& b, int color, bool mark_vitals, bool mark_pass_dead, std::uint8_t* out) {
    using detail::Components;
    const int n = b.board_size, S = b.stride, V = S * S;
    const int opp = color ^ 1;
    const int d4[4] = {-S, -1, +1, +S};
    const int d8[4] = {-S - 1, -S + 1, +S - 1, +S + 1};
    const std::uint8_t* st = b.state;

    std::uint8_t occ[kMaxVertices];        // `color` where a stone of that colour stands, empty elsewhere on the board
    for (int v = 0; v < V; ++v) occ[v] = st[v] == kInvalid ? kInvalid : (st[v] == color ? color : kEmpty);

    Components regions, strings;
    regions.Build(occ, kEmpty, n, S);
    strings.Build(occ, color, n, S);
    const int nr = regions.count, ns = strings.count;

    // ---- potential vital regions, and for each the strings it is healthy for (at most four) ---------------------------
    bool vital[kMaxVertices / 2 + 2];
    bool vacuous[kMaxVertices / 2 + 2];          // no empty point at all: healthy for every adjacent string
    std::int16_t healthy[kMaxVertices / 2 + 2][4];
    // (a 4-connected n x n grid has at most ceil(n*n / 2) components of one feature value)
    static_assert(kMaxVertices / 2 + 2 >= (kMaxBoardSize * kMaxBoardSize + 1) / 2 + 1, "component bound");
    for (int r = 1; r <= nr; ++r) {
        bool ok = true, first = true;
        std::int16_t hs[4] = {0, 0, 0, 0};
        int nh = 0;
        for (int i = regions.begin[r]; i < regions.begin[r + 1] && ok; ++i) {
            const int p = regions.member[i];
            if (st[p] != kEmpty) continue;               // an opposing stone: always fine (board.cc:1769-1772)
            std::int16_t adj[4];
            int na = 0;
            for (int k = 0; k < 4; ++k) {
                const int s = strings.label[p + d4[k]];
                if (!s) continue;
                bool seen = false;
                for (int j = 0; j < na; ++j) seen |= adj[j] == s;
                if (!seen) adj[na++] = (std::int16_t)s;
            }
            if (na == 0) {                               // an empty point that touches no stone of `color`
                ok = false;
                break;
            }
            if (first) {
                for (int j = 0; j < na; ++j) hs[j] = adj[j];
                nh = na;
                first = false;
            } else {
                int keep = 0;
                for (int j = 0; j < nh; ++j) {
                    bool in = false;
                    for (int q = 0; q < na; ++q) in |= adj[q] == hs[j];
                    if (in) hs[keep++] = hs[j];
                }
                nh = keep;
            }
        }
        vital[r] = ok;
        vacuous[r] = ok && first;
        for (int j = 0; j < 4; ++j) healthy[r][j] = j < nh ? hs[j] : 0;
    }

    // ---- Benson: drop strings with fewer than two healthy vital regions until nothing changes ----------------------
    bool alive[kMaxVertices / 2 + 2];
    std::int16_t stamp[kMaxVertices / 2 + 2];
    for (int s = 1; s <= ns; ++s) alive[s] = true;
    for (int r = 0; r <= nr; ++r) stamp[r] = 0;
    int epoch = 0;
    for (bool changed = true; changed;) {
        changed = false;
        for (int s = 1; s <= ns; ++s) {
            if (!alive[s]) continue;
            ++epoch;
            if (epoch == 32767) {                        // cannot happen on boards this small; keep the stamps sound anyway
                for (int r = 0; r <= nr; ++r) stamp[r] = 0;
                epoch = 1;
            }
            int good = 0;
            for (int i = strings.begin[s]; i < strings.begin[s + 1] && good < 2; ++i) {
                const int p = strings.member[i];
                for (int k = 0; k < 4; ++k) {
                    const int r = regions.label[p + d4[k]];
                    if (!r || !vital[r] || stamp[r] == epoch) continue;
                    stamp[r] = (std::int16_t)epoch;
                    bool mine = vacuous[r];
                    for (int j = 0; j < 4; ++j) mine |= healthy[r][j] == s;
                    good += mine;
                }
            }
            if (good >= 2) continue;
            alive[s] = false;
            changed = true;
            for (int i = strings.begin[s]; i < strings.begin[s + 1]; ++i) {
                const int p = strings.member[i];
                occ[p] = kEmpty;
                for (int k = 0; k < 4; ++k) vital[regions.label[p + d4[k]]] = false;   // label 0 = scratch entry
            }
        }
    }

    auto mark = [&](int v) { out[(v / S - 1) * n + (v % S - 1)] = 1; };
    for (int s = 1; s <= ns; ++s)
        if (alive[s])
            for (int i = strings.begin[s]; i < strings.begin[s + 1]; ++i) mark(strings.member[i]);
    if (mark_vitals) {
        for (int r = 1; r <= nr; ++r) {
            if (!vital[r]) continue;
            for (int i = regions.begin[r]; i < regions.begin[r + 1]; ++i) {
                mark(regions.member[i]);
                occ[regions.member[i]] = (std::uint8_t)color;
            }
        }
    }
    if (!mark_pass_dead) return;

    // ---- pass-dead regions of the opponent ---------------------------------------------------------------------------
    regions.Build(occ, kEmpty, n, S);
    Components rest;                                  // built on demand: the board minus the scanned region
    std::uint8_t inside[kMaxVertices];
    bool inner[kMaxVertices];
    for (int r = 1; r <= regions.count; ++r) {
        const int lo = regions.begin[r], hi = regions.begin[r + 1];
        bool have_inner = false;
        int eyes = 0, eye_at[2] = {0, 0};
        for (int i = lo; i < hi && eyes < 3; ++i) {
            const int p = regions.member[i];
            if (st[p] == opp) continue;                // its own stone cannot become its eye (board.cc:1971-1974)
            bool side_taken = false;
            for (int k = 0; k < 4; ++k) side_taken |= occ[p + d4[k]] == color;
            if (side_taken) continue;
            int corner_color = 0, corner_off = 0, corner_color_raw = 0;
            for (int k = 0; k < 4; ++k) {
                const int f = occ[p + d8[k]];
                corner_off += f == kInvalid;
                corner_color_raw += f == color;
            }
            if (corner_color_raw > (corner_off ? 0 : 1)) {
                // only now can an inner region change the verdict (it turns a `color` corner into the owner's)
                if (!have_inner) {
                    for (int v = 0; v < V; ++v) inside[v] = st[v] == kInvalid ? kInvalid : kEmpty;
                    for (int j = lo; j < hi; ++j) inside[regions.member[j]] = (std::uint8_t)color;
                    rest.Build(inside, kEmpty, n, S);
                    std::memset(inner, 0, sizeof(inner[0]) * V);
                    // erase-while-iterating of the reference: the component after an edge-touching one is not examined
                    for (int c = 1; c <= rest.count; ++c) {
                        bool touches_edge = false;
                        for (int j = rest.begin[c]; j < rest.begin[c + 1] && !touches_edge; ++j) {
                            const int v = rest.member[j];
                            for (int k = 0; k < 4; ++k) touches_edge |= inside[v + d4[k]] == kInvalid;
                        }
                        if (!touches_edge) {
                            for (int j = rest.begin[c]; j < rest.begin[c + 1]; ++j) inner[rest.member[j]] = true;
                        } else if (c + 1 <= rest.count) {
                            ++c;                          // skipped by the reference: stays in its list = inner
                            for (int j = rest.begin[c]; j < rest.begin[c + 1]; ++j) inner[rest.member[j]] = true;
                        }
                    }
                    have_inner = true;
                }
                for (int k = 0; k < 4; ++k) {
                    const int a = p + d8[k];
                    corner_color += !inner[a] && occ[a] == color;
                }
            } else {
                corner_color = corner_color_raw;
            }
            if (corner_color > (corner_off ? 0 : 1)) continue;
            if (eyes < 2) eye_at[eyes] = p;
            ++eyes;
        }
        if (eyes == 2) {
            const int delta = eye_at[1] - eye_at[0];
            if (delta == 1 || delta == -1 || delta == S || delta == -S) eyes = 1;
        }
        if (eyes < 2)
            for (int i = lo; i < hi; ++i) mark(regions.member[i]);
    }
}

// Board::ComputeReachArea (/root/reference/src/game/board.cc:1547-1579, flood fills :309-349): out[y * board_size + x]
// = kBlack / kWhite when the point is a stone of that colour or is reached from such stones through empty points and
// is NOT reached that way by the other colour; kEmpty otherwise (reached by both or by none).  Second half of
// Board::ComputeScoreArea, which the encoder asks for at every leaf.
inline void ReachArea(const BoardView& b, int* out) {
    const int n = b.board_size, S = b.stride;
    const int d4[4] = {-S, -1, +1, +S};
    const std::uint8_t* st = b.state;
    std::uint8_t reached[kMaxVertices];          // bit 0: from black, bit 1: from white
    std::int16_t queue[kMaxVertices];
    std::memset(reached, 0, (size_t)S * S);
    for (int color = 0; color < 2; ++color) {
        const std::uint8_t bit = (std::uint8_t)(1 << color);
        int head = 0, tail = 0;
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                const int v = (y + 1) * S + x + 1;
                if (st[v] == color) {
                    reached[v] |= bit;
                    queue[tail++] = (std::int16_t)v;
                }
            }
        while (head < tail) {
            const int v = queue[head++];
            for (int k = 0; k < 4; ++k) {
                const int a = v + d4[k];
                if (st[a] == kEmpty && !(reached[a] & bit)) {
                    reached[a] |= bit;
                    queue[tail++] = (std::int16_t)a;
                }
            }
        }
    }
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) {
            const std::uint8_t r = reached[(y + 1) * S + x + 1];
            out[y * n + x] = r == 1 ? kBlack : r == 2 ? kWhite : kEmpty;
        }
}

}  // namespace sb_go
