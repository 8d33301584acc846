// Ladder map of a Go position — what the reference computes in Board::GetLadderMap
// (/root/reference/src/game/board.cc:1618-1688) through Board::IsLadder / PreyMove / HunterMove / PreySelections /
// HunterSelections (:520-817) — SURVEY.md §8(f) rank 2 ("Encoder on the critical path ... ladder map"): four of the 43
// input planes (Encoder::FillLadder, encoder.cc:248-266), recomputed for EVERY network evaluation.
//
// Why it is here: after the pass-alive replacement (pass_alive.h) ladder reading is the largest remaining piece of the
// encoder (gprof of the host loop, DESIGN.md §5b: ~100 search nodes per evaluated position).  The reference reads
// ladders on full `Board` objects: every fork is `new Board(*board)` (~10 KB with Zobrist keys, empty-point lists and
// prisoner counts that a ladder never looks at), every node allocates std::vectors for its candidate moves, and every
// stone played maintains hashes and lists.  This restatement runs the SAME search — same candidate moves in the same
// order, same 2000-node budget, same string bookkeeping — on a 4 KB plain struct with fixed arrays.
//
// "Same" has to include the order in which liberties of a string are found, because that is the order in which
// candidate moves are tried and the node budget is spent: the walk follows the circular `next` list of the string from
// the stone the scan met first, and that list is the product of every merge since the stones were played
// (Board::MergeStrings, board.cc:1345-1375, splices by swapping two `next` entries).  So the search state is a copy of
// the reference's own string arrays (next / parent / liberties / stones, game/strings.h:7-19, including the sentinel
// entry that off-board and empty neighbours point to) and its 4-bit neighbour counters (board.h:258), and every update
// below restates the reference's update of exactly those arrays (AddStone :1281, RemoveStone :1313, MergeStrings :1345,
// RemoveString :1377, UpdateBoard :1407, PlayMoveAssumeLegal :1484 — minus hashes, empty lists, prisoners, passes).
// The parity harness (oracle/pass_alive_harness.cc, modes `ladder` / `dumpladder`) compares Board::GetLadderMap of the
// unmodified reference with this code on every position of seeded random games; tests/test_pass_alive.py replays
// committed fixtures without the reference.
//
// Header-only, no dependency on the reference.  Vertex coding as in pass_alive.h: 0 black, 1 white, 2 empty,
// 3 off-board; vertex = (y + 1) * stride + x + 1 with stride = board_size + 2.
#pragma once

#include <cstdint>
#include <cstring>

namespace sb_go {

constexpr int kLadderMaxBoard = 19;                                       // kBoardSize, game/types.h:5-12
constexpr int kLadderVertices = (kLadderMaxBoard + 2) * (kLadderMaxBoard + 2);   // kNumVertices
constexpr int kLadderSentinel = kLadderVertices;                          // Strings: parent of every non-stone vertex
constexpr int kLadderMaxNodes = 2000;                                     // kMaxLadderNodes, game/types.h:68

// LadderType, game/types.h:70-82 (values as the reference's enum)
enum LadderResult : int { kGoodForHunter = 0, kGoodForPrey, kGoodForNeither, kLadderDeath, kLadderEscapable, kLadderAtari, kLadderTake, kNotLadder };

struct LadderBoard {
    std::uint8_t state[kLadderVertices];
    std::uint16_t neighbours[kLadderVertices];     // 4-bit counters: black | white << 4 | empty << 8 (types.h:29-41)
    std::uint16_t next[kLadderVertices + 1];
    std::uint16_t parent[kLadderVertices + 1];
    std::uint16_t liberties[kLadderVertices + 1];
    std::uint16_t stones[kLadderVertices + 1];
    int ko_move;                                   // kNullVertex (0) when none
    int board_size;
    int stride;
    int dir[4];
};

namespace ladder_detail {

constexpr int kBlackL = 0, kWhiteL = 1, kEmptyL = 2;
constexpr int kNullVertexL = 0;

// candidate moves of one node.  128 entries cannot overflow on a 19 x 19 board: every entry beyond the first is the single
// liberty of a DIFFERENT enemy string in atari next to the prey (>= 2 vertices each, plus a prey stone to touch it)
struct MoveList {
    static constexpr int kCap = 128;
    std::int16_t v[kCap];
    int n = 0;
    bool Has(int x) const {
        for (int i = 0; i < n; ++i)
            if (v[i] == x) return true;
        return false;
    }
    void Push(int x) {
        if (n < kCap) v[n++] = (std::int16_t)x;
    }
};

inline int Plibs(const LadderBoard& b, int vtx) { return (b.neighbours[vtx] >> 8) & 0xf; }     // CountPliberties, board.cc:387
inline bool SimpleEye(const LadderBoard& b, int vtx, int color) { return (b.neighbours[vtx] & (4 << (4 * color))) != 0; }   // :901

// Board::FindStringLiberties (board.cc:416-436): liberties of the string in the order the walk from `vtx` meets them,
// appended to `buf` unless already there; returns how many were appended.
inline int FindLiberties(const LadderBoard& b, int vtx, MoveList& buf) {
    int found = 0, next = vtx;
    do {
        for (int k = 0; k < 4; ++k) {
            const int a = next + b.dir[k];
            if (b.state[a] == kEmptyL && !buf.Has(a)) {
                buf.Push(a);
                ++found;
            }
        }
        next = b.next[next];
    } while (next != vtx);
    return found;
}

// Board::FindStringLibertiesGainingCaptures (board.cc:438-467).  As written there the "already seen" list stays empty, so
// every adjacent enemy stone whose string is in atari contributes that string's liberty (if new).
inline int FindCaptureLiberties(const LadderBoard& b, int vtx, MoveList& buf) {
    const int opp = !b.state[vtx];
    int found = 0, next = vtx;
    do {
        for (int k = 0; k < 4; ++k) {
            const int a = next + b.dir[k];
            if (b.state[a] == opp && b.liberties[b.parent[a]] == 1) found += FindLiberties(b, a, buf);
        }
        next = b.next[next];
    } while (next != vtx);
    return found;
}

// Board::IsSuicide (board.cc:940-960)
inline bool IsSuicide(const LadderBoard& b, int vtx, int color) {
    if (Plibs(b, vtx)) return false;
    for (int k = 0; k < 4; ++k) {
        const int a = vtx + b.dir[k];
        const int libs = b.liberties[b.parent[a]];
        const int st = b.state[a];
        if (st == color && libs > 1) return false;
        if (st == !color && libs <= 1) return false;
    }
    return true;
}

// Board::IsLegalMove(vtx, color) for an on-board vertex (board.cc:203-231): empty, not suicide, not the ko point
inline bool IsLegal(const LadderBoard& b, int vtx, int color) {
    return b.state[vtx] == kEmptyL && !IsSuicide(b, vtx, color) && vtx != b.ko_move;
}

// Board::GetLadderLiberties (board.cc:469-518)
inline void LadderLiberties(const LadderBoard& b, int vtx, int color, int& lower, int& upper) {
    const int stone_libs = Plibs(b, vtx);
    const int opp = !color;
    int num_captures = 0, potential = 0, num_connection = 0, max_connection = stone_libs;
    for (int k = 0; k < 4; ++k) {
        const int a = vtx + b.dir[k];
        const int st = b.state[a];
        if (st == color) {
            const int alibs = b.liberties[b.parent[a]] - 1;
            num_connection += alibs;
            if (alibs > max_connection) max_connection = alibs;
        } else if (st == opp) {
            const int aip = b.parent[a];
            if (b.liberties[aip] == 1) {
                ++num_captures;
                potential += b.stones[aip];
            }
        }
    }
    lower = num_captures + max_connection;
    upper = stone_libs + potential + num_connection;
}

// ---- playing a stone: Board::UpdateBoard (board.cc:1407-1464) on the arrays a ladder reads -----------------------
inline void AddStone(LadderBoard& b, int vtx, int color) {                     // board.cc:1281-1311
    int pars[4], n = 0;
    b.state[vtx] = (std::uint8_t)color;
    for (int k = 0; k < 4; ++k) {
        const int a = vtx + b.dir[k];
        b.neighbours[a] = (std::uint16_t)(b.neighbours[a] + (1u << (4 * color)) - (1u << 8));
        const int ip = b.parent[a];
        bool seen = false;
        for (int i = 0; i < n; ++i) seen |= pars[i] == ip;
        if (!seen) {
            b.liberties[ip]--;
            pars[n++] = ip;
        }
    }
}
inline void RemoveStone(LadderBoard& b, int vtx, int color) {                  // board.cc:1313-1343
    int pars[4], n = 0;
    b.state[vtx] = kEmptyL;
    for (int k = 0; k < 4; ++k) {
        const int a = vtx + b.dir[k];
        b.neighbours[a] = (std::uint16_t)(b.neighbours[a] + (1u << 8) - (1u << (4 * color)));
        const int ip = b.parent[a];
        bool seen = false;
        for (int i = 0; i < n; ++i) seen |= pars[i] == ip;
        if (!seen) {
            b.liberties[ip]++;
            pars[n++] = ip;
        }
    }
}
inline void MergeStrings(LadderBoard& b, int ip, int aip) {                    // board.cc:1345-1375
    b.stones[ip] = (std::uint16_t)(b.stones[ip] + b.stones[aip]);
    int pos = aip;
    do {
        for (int k = 0; k < 4; ++k) {
            const int a = pos + b.dir[k];
            if (b.state[a] == kEmptyL) {
                bool found = false;
                for (int kk = 0; kk < 4; ++kk) {
                    if (b.parent[a + b.dir[kk]] == ip) {
                        found = true;
                        break;
                    }
                }
                if (!found) b.liberties[ip]++;
            }
        }
        b.parent[pos] = (std::uint16_t)ip;
        pos = b.next[pos];
    } while (pos != aip);
    const std::uint16_t t = b.next[aip];
    b.next[aip] = b.next[ip];
    b.next[ip] = t;
}
inline int RemoveString(LadderBoard& b, int ip) {                              // board.cc:1377-1398
    int pos = ip, removed = 0;
    const int color = b.state[ip];
    do {
        RemoveStone(b, pos, color);
        b.parent[pos] = kLadderSentinel;
        ++removed;
        pos = b.next[pos];
    } while (pos != ip);
    return removed;
}
// Board::PlayMoveAssumeLegal for a stone (board.cc:1484-1507): returns nothing, sets the ko point
inline void Play(LadderBoard& b, int vtx, int color) {
    AddStone(b, vtx, color);
    b.next[vtx] = (std::uint16_t)vtx;                                          // Strings::AddStone, strings.h:49-54
    b.parent[vtx] = (std::uint16_t)vtx;
    b.liberties[vtx] = (std::uint16_t)Plibs(b, vtx);
    b.stones[vtx] = 1;
    const bool is_eyeplay = SimpleEye(b, vtx, !color);
    int captured = 0, captured_vtx = kNullVertexL;
    for (int k = 0; k < 4; ++k) {
        const int a = vtx + b.dir[k];
        const int aip = b.parent[a];
        const int st = b.state[a];
        if (st == !color) {
            if (b.liberties[aip] == 0) {                    // "<= 0" on an unsigned counter in the reference
                captured += RemoveString(b, a);
                captured_vtx = a;
            }
        } else if (st == color) {
            const int ip = b.parent[vtx];
            if (ip != aip) {
                if (b.stones[ip] >= b.stones[aip]) MergeStrings(b, ip, aip);
                else MergeStrings(b, aip, ip);
            }
        }
    }
    if (b.liberties[b.parent[vtx]] == 0) RemoveString(b, vtx);                 // suicide: never chosen by a ladder, kept for identity
    b.ko_move = (captured == 1 && is_eyeplay) ? captured_vtx : kNullVertexL;
}

// ---- the search (board.cc:520-773) ---------------------------------------------------------------------------
inline int PreySelections(const LadderBoard& b, int prey_color, int ladder_vtx, MoveList& sel, bool think_ko) {
    const int libs = b.liberties[b.parent[ladder_vtx]];
    if (libs >= 2 || (b.ko_move != kNullVertexL && think_ko)) return kGoodForPrey;
    FindLiberties(b, ladder_vtx, sel);
    const int not_cap_move = sel.v[0];
    FindCaptureLiberties(b, ladder_vtx, sel);
    int n = 0;
    for (int i = 0; i < sel.n; ++i)
        if (IsLegal(b, sel.v[i], prey_color)) sel.v[n++] = sel.v[i];
    sel.n = n;
    if (n == 0) return kGoodForHunter;
    if (sel.Has(not_cap_move)) {
        int lower, upper;
        LadderLiberties(b, not_cap_move, prey_color, lower, upper);
        if (lower >= 3) return kGoodForPrey;
        if (n == 1 && upper == 1) return kGoodForHunter;
    }
    return kGoodForNeither;
}

inline int HunterSelections(const LadderBoard& b, int prey_color, int ladder_vtx, MoveList& sel) {
    const int libs = b.liberties[b.parent[ladder_vtx]];
    if (libs >= 3) return kGoodForPrey;
    if (libs <= 1) return kGoodForHunter;
    MoveList buf;
    FindLiberties(b, ladder_vtx, buf);
    const int m1 = buf.v[0], m2 = buf.v[1];
    bool neighbor = false;
    for (int k = 0; k < 4; ++k) neighbor |= (m1 + b.dir[k]) == m2;
    if (!neighbor) {
        const int hunter = !prey_color;
        const int l1 = Plibs(b, m1), l2 = Plibs(b, m2);
        if (l1 >= 3 && l2 >= 3) return kGoodForPrey;
        if (l1 >= 3) {
            if (IsLegal(b, m1, hunter)) sel.Push(m1);
        } else if (l2 >= 3) {
            if (IsLegal(b, m2, hunter)) sel.Push(m2);
        } else {
            if (IsLegal(b, m1, hunter)) sel.Push(m1);
            if (IsLegal(b, m2, hunter)) sel.Push(m2);
        }
    } else {
        sel.Push(m1);
        sel.Push(m2);
    }
    return sel.n == 0 ? kGoodForPrey : kGoodForNeither;
}

inline int HunterMove(LadderBoard& board, int prey_vtx, int prey_color, int ladder_vtx, int& nodes, bool fork);

// `fork`: the caller still needs its board (it has more than one candidate): work on a copy, as the reference does
inline int PreyMove(LadderBoard& board, int hunter_vtx, int prey_color, int ladder_vtx, int& nodes, bool fork) {
    if (++nodes >= kLadderMaxNodes) return kGoodForPrey;
    LadderBoard copy;
    LadderBoard* b = &board;
    if (fork) {
        std::memcpy(&copy, &board, sizeof(LadderBoard));
        b = &copy;
    }
    if (hunter_vtx != kNullVertexL) Play(*b, hunter_vtx, !prey_color);
    MoveList sel;
    const int res = PreySelections(*b, prey_color, ladder_vtx, sel, hunter_vtx != kNullVertexL);
    if (res != kGoodForNeither) return res;
    const bool next_fork = sel.n != 1;
    int best = kGoodForNeither;
    for (int i = 0; i < sel.n; ++i) {
        best = HunterMove(*b, sel.v[i], prey_color, ladder_vtx, nodes, next_fork);
        if (best == kGoodForPrey) break;
    }
    return best;
}

inline int HunterMove(LadderBoard& board, int prey_vtx, int prey_color, int ladder_vtx, int& nodes, bool fork) {
    if (++nodes >= kLadderMaxNodes) return kGoodForPrey;
    LadderBoard copy;
    LadderBoard* b = &board;
    if (fork) {
        std::memcpy(&copy, &board, sizeof(LadderBoard));
        b = &copy;
    }
    if (prey_vtx != kNullVertexL) Play(*b, prey_vtx, prey_color);
    MoveList sel;
    const int res = HunterSelections(*b, prey_color, ladder_vtx, sel);
    if (res != kGoodForNeither) return res;
    const bool next_fork = sel.n != 1;
    int best = kGoodForNeither;
    for (int i = 0; i < sel.n; ++i) {
        best = PreyMove(*b, sel.v[i], prey_color, ladder_vtx, nodes, next_fork);
        if (best == kGoodForHunter) break;
    }
    return best;
}

// Board::IsLadder (board.cc:775-817): vital moves = the hunter's moves that win the ladder
inline bool IsLadder(const LadderBoard& base, int vtx, MoveList& vital) {
    const int prey_color = base.state[vtx];
    vital.n = 0;
    // strings_.liberties_ is the exact number of distinct liberties (AddStone / RemoveStone / MergeStrings keep it so; the
    // reference asserts it in HunterSelections, board.cc:593): three or more is no ladder, without walking the string
    if (base.liberties[base.parent[vtx]] >= 3) return false;
    MoveList buf;
    const int libs = FindLiberties(base, vtx, buf);
    int nodes = 0;
    LadderBoard work;
    if (libs == 1) {
        std::memcpy(&work, &base, sizeof(LadderBoard));
        if (PreyMove(work, kNullVertexL, prey_color, vtx, nodes, false) == kGoodForHunter) vital.Push(buf.v[0]);
    } else if (libs == 2) {
        for (int i = 0; i < 2; ++i) {
            const int v = buf.v[i];
            std::memcpy(&work, &base, sizeof(LadderBoard));
            if (IsLegal(work, v, !prey_color) && PreyMove(work, v, prey_color, vtx, nodes, false) == kGoodForHunter) vital.Push(v);
        }
    }
    return vital.n != 0;
}

}  // namespace ladder_detail

// Board::GetLadderMap (board.cc:1618-1688): out[y * board_size + x] = LadderResult (kNotLadder, kLadderDeath,
// kLadderEscapable on the stones of a string that loses a ladder; kLadderTake / kLadderAtari on the hunter's winning moves).
inline void LadderMap(const LadderBoard& b, std::uint8_t* out) {
    using namespace ladder_detail;
    const int n = b.board_size;
    for (int i = 0; i < n * n; ++i) out[i] = (std::uint8_t)kNotLadder;
    // one verdict per string, keyed by its parent vertex (the reference keeps two lists and searches them linearly)
    std::uint8_t verdict[kLadderVertices + 1];     // 0 unknown, 1 ladder, 2 not a ladder
    std::memset(verdict, 0, sizeof(verdict));
    for (int y = 0; y < n; ++y) {
        for (int x = 0; x < n; ++x) {
            const int idx = y * n + x;
            const int vtx = (y + 1) * b.stride + x + 1;
            if (b.state[vtx] == kEmptyL) continue;
            const int parent = b.parent[vtx];
            MoveList vital;
            bool first_found = false;
            if (verdict[parent] == 0) {
                if (IsLadder(b, vtx, vital)) {
                    verdict[parent] = 1;
                    first_found = true;
                } else {
                    verdict[parent] = 2;
                    continue;
                }
            } else if (verdict[parent] == 2) {
                continue;
            }
            const int libs = b.liberties[parent];
            out[idx] = (std::uint8_t)(libs == 1 ? kLadderDeath : kLadderEscapable);
            if (first_found) {
                for (int i = 0; i < vital.n; ++i) {
                    const int v = vital.v[i];
                    const int ax = v % b.stride - 1, ay = v / b.stride - 1;
                    out[ay * n + ax] = (std::uint8_t)(libs == 1 ? kLadderTake : kLadderAtari);
                }
            }
        }
    }
}

}  // namespace sb_go
