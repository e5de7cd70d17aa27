// flat_game.hpp — small extensive-form games as flat device tables.
//
// The reference drives MCCFR through the `CfrGame`/`CfrEncoder`/`CfrInfo` plug-in traits
// (crates/mccfr/src/state/game.rs:21-53, strategy/encoder.rs:21-90): `turn`, `apply`, `payoff`, `info`,
// `choices`, `branches`.  A GPU cannot call back into Rust per node, so a small game is enumerated ONCE
// on the host into breadth-first arrays (children contiguous, depth levels contiguous) and every kernel
// walks those arrays.  A sampled MCCFR tree is a sub-tree of this enumeration; the exploitability tree
// (solver.rs:327-338, VanillaSampling from `exploitability_root()`) is the enumeration itself.
#pragma once
#include <cstdint>
#include <vector>

namespace rbp {

enum : uint8_t { TURN_P0 = 0, TURN_P1 = 1, TURN_CHANCE = 2, TURN_TERMINAL = 3 };
constexpr int kMaxActions = 4;  // widest decision node of a flat game (Kuhn/Leduc: 2)

struct FlatNode {         // 16 B, one LDG.128 per node
    int32_t first_child;  // children are [first_child, first_child + n_child), in `branches()` order
    uint32_t info_key;    // packed CfrInfo (RNG-contract word, include/rbp.h)
    int16_t info;         // dense infoset index for player nodes, -1 otherwise
    uint8_t turn;
    uint8_t n_child;
    float payoff0;        // terminal utility of player 0 (player 1 in payoff1[])
};

struct FlatGame {
    std::vector<FlatNode> nodes;
    std::vector<float> payoff1;
    std::vector<int32_t> parent;
    std::vector<int32_t> lifo_index;  // node index the reference's LIFO TreeBuilder (builder.rs:141-160) assigns
    std::vector<int32_t> level_start; // BFS depth levels: nodes of depth d are [level_start[d], level_start[d+1])
    // per infoset
    std::vector<uint32_t> info_key;
    std::vector<uint8_t> info_player, info_actions;
    std::vector<int32_t> info_row;    // first row of the infoset in the Encounter table
    std::vector<int32_t> span_start;  // CSR over nodes of each infoset, sorted by lifo_index
    std::vector<int32_t> span_nodes;
    int n_rows = 0;
    // root rule of `CfrGame::root()`: two-swap Fisher-Yates over `deck` cards (kuhn/leduc game.rs root())
    int deck = 0;
    std::vector<int32_t> root_table;  // [c0 * deck + c1] -> node id (or -1)
    // sizing (host DP over the enumeration)
    int max_tree_nodes = 0, max_tree_infos = 0, max_depth = 0, n_terminals = 0;
};

// game ids follow include/rbp.h (RBP_GAME_KUHN = 0, RBP_GAME_LEDUC = 1)
bool build_flat_game(int game_id, FlatGame* out);

}  // namespace rbp
