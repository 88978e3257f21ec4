/*
 * nz_engine.h — C ABI of the B200 self-play search engine (libnz_engine.so).
 *
 * The reference (guilherme439/NuZero) has no FFI: its boundary for this path is Python duck typing.
 * Each entry point below names the reference interface it stands in for (paths relative to the
 * reference root).  The Python mirror of that interface lives in nuzero_b200/ and binds these
 * symbols with ctypes (see INTEGRATION.md for the stub a NuZero maintainer would add).
 *
 * Conventions
 *   - plain C types only; every device pointer is allocated and owned by the caller (PyTorch);
 *     the library never allocates or frees device memory and keeps no host pointer past a call;
 *   - every call is asynchronous on the given stream (a cudaStream_t passed as void*), performs no
 *     synchronisation and no allocation, and is therefore CUDA-graph capturable;
 *   - return value 0 = ok, <0 = error; nz_last_error() gives the message (thread local).  The
 *     reference raises Python exceptions at the same places (Games/SCS/SCS_Game.py:382,480,...);
 *     the Python wrapper re-raises.  Device-side faults (node pool exhausted, illegal action)
 *     never trap: they set a per-game error word that the host reads at move boundaries;
 *   - one handle per GPU, not thread-safe per handle.
 */
#ifndef NZ_ENGINE_H
#define NZ_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NZ_ABI_VERSION 3

enum { NZ_GAME_TTT = 0, NZ_GAME_SCS = 1 };
enum { NZ_F32 = 0, NZ_BF16 = 1 };

/* per-game phase (ctl word), see nz_engine_buffer("ctl") */
enum {
  NZ_PHASE_READY = 0,        /* may run simulations */
  NZ_PHASE_LEAF_PENDING = 1, /* a leaf row was emitted; waits for the network's policy/value row */
  NZ_PHASE_MOVE_READY = 2,   /* manual mode: all simulations of the move done, waits for nz_commit_moves */
  NZ_PHASE_IDLE = 3,         /* quota of games played, or game over in manual mode */
  NZ_PHASE_ERROR = 4,
  NZ_PHASE_DESCENDING = 5    /* a descent used up the launch's level budget; it resumes in the next launch */
};

/* per-game error bits */
enum {
  NZ_ERR_POOL_FULL = 1,   /* node pool exhausted */
  NZ_ERR_DEPTH = 2,       /* search path longer than max_depth */
  NZ_ERR_ILLEGAL = 4,     /* commit of an action that is not a child of the root */
  NZ_ERR_ARENA_FULL = 8,  /* move-record arena full; record dropped */
  NZ_ERR_CTABLE = 16      /* visit count beyond the host log table; device log() used */
};

/*
 * Search + game configuration.  Field names follow the reference's search-config YAML
 * (Configs/Search/Examples/documentation_search_config.yaml) where one exists.
 */
typedef struct nz_config {
  int32_t abi_version;       /* NZ_ABI_VERSION */
  int32_t game_kind;         /* NZ_GAME_TTT | NZ_GAME_SCS */
  int32_t n_games;           /* G: concurrent game slots on this GPU */
  int32_t pool_nodes;        /* node-pool capacity per slot */
  int32_t max_depth;         /* longest root->leaf path storable */
  int32_t max_children;      /* widest expansion (sizes the noised-root side array) */
  int32_t mcts_simulations;  /* Simulation.mcts_simulations   (Search/Explorer.py:48) */
  int32_t training;          /* Explorer(search_config, training) (Explorer.py:35-37) */
  int32_t policy_is_prob;    /* 1: network emits probabilities (parity stub); 0: logits -> softmax (Explorer.py:152,159) */
  int32_t leaf_dtype;        /* NZ_F32 | NZ_BF16: element type of the leaf tensor */
  int32_t policy_dtype;      /* NZ_F32 | NZ_BF16: element type of the policy rows */
  int32_t auto_advance;      /* 1: engine commits moves / restarts games itself (batched Gamer); 0: manual (Explorer.run_mcts) */
  int32_t games_per_slot;    /* auto mode: games each slot plays before going idle; 0 = no limit */
  int32_t max_sims_per_launch; /* simulations one game may run inside one nz_advance launch */
  int32_t record_detail;     /* 1: move records also carry child W and priors (parity tests) */
  int32_t number_of_softmax_moves;   /* Exploration.* (Explorer.py:74-89) */
  double pb_c_base;          /* UCT.pb_c_base (Explorer.py:105) */
  double pb_c_init;          /* UCT.pb_c_init (Explorer.py:106) */
  double value_factor;       /* Exploration.value_factor (Explorer.py:122) */
  double root_exploration_fraction; /* Explorer.py:203 */
  double root_dist_alpha;    /* Explorer.py:204 */
  double root_dist_beta;     /* Explorer.py:205 */
  double epsilon_softmax_exploration; /* Explorer.py:79 */
  double epsilon_random_exploration;  /* Explorer.py:80 */
  uint64_t seed;             /* Philox key for device-drawn noise (throughput mode) */
  int32_t ctable_len;        /* entries of the host-computed exploration-bias table c[N] */
  int32_t tape_moves;        /* >0: replay pre-drawn random numbers (parity), rows per slot */
  int32_t tape_width;        /* gamma draws per tape row */
  int32_t arena_words;       /* capacity of the move-record arena in 32-bit words */
  /* SCS only: host pointer to a flat int32 description of the scenario (see nuzero_b200/games/scs.py),
   * read during nz_engine_create and not retained. */
  const int32_t* scs_desc;
  int32_t scs_desc_len;
  int32_t compact_on_reroot; /* 1 (auto mode): a slot's pool is two halves; when the current half may not hold another move's
                              * growth, the committed child's live sub-tree is copied breadth first into the other half */
  int32_t max_levels_per_launch; /* tree levels one game may descend inside one launch (0 = no limit): bounds the
                                  * launch time when a few games have very deep trees */
  int32_t virtual_loss_width; /* V: simulations one game may have in flight at the network (0 / 1 = one, the reference's
                              * sequencing, bit-exact).  V > 1 is a throughput mode, NOT result-identical to the reference: a
                              * descent that ends in a fresh non-terminal leaf adds a virtual visit (N += 1) to its path and the
                              * game starts another descent; the leaf tensor / policy / value then hold G * V rows, row g * V + j
                              * for the j-th pending leaf of game g.  A descent that reaches a leaf already waiting for the network
                              * ends the game's launch without effect. */
  int32_t node_state_cache;  /* 1 (SCS): every expanded node keeps its compact game state ("nstate" buffer, pool_nodes / 2 rows
                              * per slot, keyed by the node's child run), so a simulation selects its way down the tree on node records alone and steps the game
                              * ONCE, from the leaf's parent, instead of replaying the whole path from the root
                              * (Explorer.py:54-58 steps a scratch game at every level).  Same results: the game is
                              * deterministic, a node's state is a function of its path.  0: replay the path (Tic-Tac-Toe
                              * always does: its state is one register). */
} nz_config;

typedef struct nz_engine nz_engine;

/* last error message of the calling thread ("" if none) */
const char* nz_last_error(void);
int nz_abi_version(void);
/* sizeof(nz_config) as this library was compiled: lets a binding check its own struct declaration */
size_t nz_config_bytes(void);

/* Create / destroy a handle (host memory only).  Stands in for constructing
 * Explorer(search_config, training) + Gamer(...) (Search/Explorer.py:35, Training/Gamer.py:20-37). */
int nz_engine_create(const nz_config* cfg, nz_engine** out);
void nz_engine_destroy(nz_engine* eng);

/* Bytes of device workspace the caller must allocate (256-byte aligned) and then bind. */
size_t nz_engine_workspace_bytes(const nz_engine* eng);
int nz_engine_bind(nz_engine* eng, void* dev_workspace, size_t bytes);

/* Locate a named sub-buffer of the workspace (for views, uploads and tests):
 * "nodes" 32 bytes x G*P: {prior f64, W f64, N i32, flags u32 (bit 0 noised f64 prior, bits 16-31 action), first child u32,
 *  n_children u32}; per slot node 0 is the current root and nodes 1 .. n_children(root) are its children,
 * "ctl" u32[G*NZ_CTL_WORDS], "path" u32[G*max_depth],
 * "ctable" f64[ctable_len*2] rows (c(N), sqrt(N)), "gamma_tape" f64[G*tape_moves*tape_width], "unif_tape" f64[G*tape_moves*3],
 * "arena" u32[arena_words], "arena_top" u32[4] = {words used, records dropped, records written, -},
 * "rec_index" u32[] arena offset of every record in the order they were counted, "scs_static" ...,
 * "nstate" u32[G*(P/2)][round4(state_words)] compact game state of every expanded node, keyed by its child run (node_state_cache),
 * "dense_count" u32[2][4] / "dense_rows" i32[2][G] the dense leaf rows of nz_engine_attach_cache (per lane). */
int nz_engine_buffer(const nz_engine* eng, const char* name, size_t* offset, size_t* bytes);

/* Start every slot on a fresh game: Node(0) root + game_class(*game_args) (Training/Gamer.py:52,59). */
int nz_reset(nz_engine* eng, void* stream);

/*
 * One launch of the search kernel over all slots.  For every slot, in order:
 *   - if a leaf is pending, consume its policy row and value (Explorer.evaluate, Explorer.py:137-181:
 *     mask, normalise, create children) and back the value up (Explorer.backpropagate :132-135);
 *   - run simulations (Explorer.run_mcts :49-62: select_child/score :99-130 + game.step) until a
 *     non-terminal leaf needs the network — its encoded state (game.generate_network_input) is written
 *     to row g of `leaf_out` — or the launch budget / the move's simulation count is reached;
 *   - auto mode: when the move's simulations are done, choose the action (Explorer.select_action
 *     :70-97), append a move record, step the real game and re-root (Training/Gamer.py:74-79), add
 *     root noise for the next move (Explorer.add_exploration_noise :201-210), restart finished games.
 * leaf_out : [G, C, R, Cc] leaf_dtype     policy_in : [G, A] policy_dtype     value_in : [G] f32
 */
int nz_advance(nz_engine* eng, void* leaf_out, const void* policy_in, const float* value_in, void* stream);

/* Manual mode (Explorer.run_mcts drop-in): commit the move of every MOVE_READY slot.
 * actions: device int32[G], <0 = use the engine's own choice.  Equivalent to the caller's
 * game.step(action) + root_node = chosen_child (Training/Gamer.py:74-79, MctsAgent.py:28-33). */
int nz_commit_moves(nz_engine* eng, const int32_t* actions, void* stream);

/* Batched environment entry points — the Game interface (Games/Game.py:3-106) over n device-resident
 * compact states (state_words u32 each; nz_env_state_words tells how many):
 *   possible_actions -> nz_env_mask, step -> nz_env_step, generate_network_input -> nz_env_encode,
 *   is_terminal/get_terminal_value/get_current_player/get_length -> nz_env_status. */
int nz_env_state_words(const nz_engine* eng);
int nz_env_reset(nz_engine* eng, uint32_t* states, const int32_t* map_ids, int n, void* stream);
int nz_env_step(nz_engine* eng, uint32_t* states, const int32_t* map_ids, const int32_t* actions, int32_t* err_out, int n, void* stream);
int nz_env_mask(nz_engine* eng, const uint32_t* states, const int32_t* map_ids, uint8_t* mask_out /* [n, A] */, int n, void* stream);
int nz_env_encode(nz_engine* eng, const uint32_t* states, const int32_t* map_ids, void* out /* [n,C,R,Cc] */, int dtype, int n, void* stream);
int nz_env_status(nz_engine* eng, const uint32_t* states, const int32_t* map_ids, int32_t* out /* [n,4]: terminal, terminal_value, player, length */, int n, void* stream);

/* Move records -> replay tuples.  Stands in for what Training/ReplayBuffer.save_game (ReplayBuffer.py:24-36) reads from a
 * finished game: game.get_state_from_history(i) (the float32 network input of the position, Gamer.py:65-66) and the policy
 * half of game.make_target(i) (visit fractions over ALL actions, store_search_statistics, SCS_Game.py:1517-1521 /
 * tic_tac_toe.py:177-182).  `words` holds move records (layout: csrc/mcts.cuh write_record; the engine's "arena" or a copy
 * of it), offsets[i] the first word of record i, dst_rows[i] the row of states_out [*, C, R, Cc] f32 and policy_out [*, A]
 * f32 that position i is written to.  The value target (the game's terminal value) is known per game, not per record:
 * the caller fills it.  All pointers are device pointers. */
int nz_replay_decode(nz_engine* eng, const uint32_t* words, const int64_t* offsets, const int64_t* dst_rows, float* states_out,
                     float* policy_out, int n, void* stream);

/* Device inference cache — the reference's Utils/Caches (DictCache.py: state -> (action probabilities / logits, value),
 * consulted in Explorer.evaluate, Explorer.py:146-155) as an exact-key open-addressing table in HBM, for the whole batch.
 * Table (caller-allocated, zeroed once): keys u32[2^capacity_log2][state_words + 1] (compact leaf state + scenario map),
 * meta i32[2^capacity_log2], cache_policy [2^capacity_log2][A] (policy dtype of the engine), cache_value f32[..].
 * nz_cache_lookup: for every leaf row that waits for the network, copy the stored output into policy / value (hit) or append
 * the row to miss_rows (counters[0] = misses, counters[1] = hits; the caller zeroes counters).  The caller evaluates the
 * missed rows and calls nz_cache_insert with the same rows.  Dense-batch form: with leaf / leaf_stage non-null the look-up
 * also copies missed row i's leaf planes to row i of leaf_stage (the batch the network runs on), and nz_cache_insert with
 * policy_out / value_out non-null takes the network's outputs from rows 0..n-1 of policy / value (the dense batch), stores
 * them in the table AND scatters them to the engine's rows (policy_out / value_out).  With those pointers null the outputs
 * are read from the engine's own rows. */
int nz_cache_lookup(nz_engine* eng, uint32_t* keys, int32_t* meta, void* cache_policy, float* cache_value, int capacity_log2,
                    void* policy, float* value, int32_t* miss_rows, int32_t* counters, const void* leaf, void* leaf_stage,
                    void* stream);
int nz_cache_insert(nz_engine* eng, uint32_t* keys, int32_t* meta, void* cache_policy, float* cache_value, int capacity_log2,
                    const void* policy, const float* value, const int32_t* rows, int n, void* policy_out, float* value_out,
                    void* stream);

/* In-kernel form of the same cache: Explorer.evaluate consults the cache BEFORE the inference (Explorer.py:146-155), so a
 * simulation whose leaf was evaluated before needs no network round at all.  nz_engine_attach_cache hands the table to the
 * search kernel (read-only there): at a non-terminal leaf nz_advance probes it, and on a hit expands the leaf from the stored
 * row and goes on with the game's next simulation (up to max_sims_per_launch per launch).  Misses take the next free row of
 * the leaf tensor — unless another game of the same launch already sends the same state there: the first game to miss claims
 * a table entry (cache_row i32[2^capacity_log2], caller-allocated: the dense row a pending entry waits for) and the others
 * wait for its row, so one launch never evaluates a state twice (DENSE rows: "dense_count" u32[lane][4] = rows handed out by the lane's last nz_advance, "dense_rows" i32[lane][G] = game slot
 * of each row), so the network runs on rows [0, dense_count) only and writes policy / value rows with the same index;
 * nz_cache_insert_dense then stores those n rows in the table.  keys == NULL detaches.  Needs virtual_loss_width <= 1.
 * miss_target > 0: a slot starts no further simulation once the launch has handed out that many rows, i.e. the launch ends when
 * a network batch is full instead of after a fixed number of simulations per game (max_sims_per_launch stays the upper bound);
 * park_target > 0: likewise once that many games wait for the network, on a row of their own or on a shared one.
 * Results are identical to a run without the cache (a hit returns exactly what the network returned for that state). */
int nz_engine_attach_cache(nz_engine* eng, uint32_t* keys, int32_t* meta, int32_t* cache_row, const void* cache_policy,
                           const float* cache_value, int capacity_log2, int miss_target, int park_target);
int nz_cache_insert_dense(nz_engine* eng, uint32_t* keys, int32_t* meta, void* cache_policy, float* cache_value, int capacity_log2,
                          const void* policy, const float* value, int n, int lane, void* stream);
/* Published expansions (optional, after nz_engine_attach_cache): exp_meta i32[2^capacity_log2] (zeroed), exp_actions
 * u16[2^capacity_log2][width], exp_priors f64[..][width].  The legal actions and priors of a state follow from the state and the
 * network's row (Explorer.py:165-179), so the first game that expands a cached state publishes its (action, prior) list and every
 * later expansion of that state copies it instead of recomputing the legal mask and the soft-max.  Lists longer than `width` are
 * not published.  exp_meta == NULL switches it off. */
int nz_engine_attach_expansions(nz_engine* eng, int32_t* exp_meta, uint16_t* exp_actions, double* exp_priors, int width);
/* Two lanes of dense rows ("dense_count" u32[2][4], "dense_rows" i32[2][G]): the nz_advance calls that follow use `lane` (0 / 1)
 * and the leaf / policy / value tensors the caller passes for it.  A game that parked in lane L is skipped by launches of the
 * other lane, so the caller may run the network on the rows of launch k (lane k & 1) on one stream while launch k + 1 searches
 * on another: launch k + 2 must wait for that network call and its nz_cache_insert_dense(lane = k & 1).  Without this call
 * everything runs in lane 0 (search and network alternate). */
int nz_engine_set_lane(nz_engine* eng, int lane);

/* SCS only: byte image of the scenario tables (terrain, schedule, maps) that the caller uploads into
 * the "scs_static" workspace buffer after nz_engine_bind (parsed from nz_config.scs_desc;
 * SCS_Game.load_game_from_config, Games/SCS/SCS_Game.py:1570-1777). */
int nz_scs_static_image(const nz_engine* eng, void* host_out, size_t bytes);

/* Geometry of the bound game: action planes/rows/cols and state channels
 * (get_action_space_shape / get_state_shape, Games/Game.py:10-14). out[6] = A_planes,R,Cc,C,R,Cc */
int nz_game_shape(const nz_engine* eng, int32_t* out6);

/* Deterministic dyadic stub network of the parity protocol (one launch; see oracle/stubnet_np.py
 * for the checker).  leaf: [n, F] leaf_dtype; salt: int32[n] or NULL; uid: u32 read at uid[i*uid_stride]
 * or NULL (effective salt = salt[i] + uid*salt_uid_mul); policy_out [n, A]; value_out f32[n]. */
int nz_stubnet_forward(const void* leaf, int leaf_dtype, const int32_t* salt, const uint32_t* uid, int uid_stride, int salt_uid_mul,
                       int n, int n_features, int n_actions, void* policy_out, int policy_dtype,
                       float* value_out, void* stream);

/* Network-side helper (the network forward itself stays PyTorch / cuBLAS): neighbour-table im2col for the
 * 7-tap hexagonal and 9-tap orthogonal convolutions of Neural_Networks/Architectures/blocks.py.
 * x [batch, cells, channels] bf16, nbr int32 [cells, taps] (-1 = off board), out [batch, cells, taps, channels];
 * relu != 0 applies max(x, 0) on the way.  channels % 8 == 0. */
int nz_im2col_bf16(const void* x, const int32_t* nbr, void* out, int batch, int cells, int taps, int channels, int relu,
                   void* stream);

/* Fused form of the same convolution (no im2col matrix in memory): out = act(gather(x) . wt^T (+ residual)) with
 * tcgen05.mma / TMEM accumulators.  x [rows, cin] bf16 (rows = batch * cells), nbr int32 [cells, taps],
 * wt [n_pad, taps * cin] bf16 (W^T, K contiguous), residual / out [rows, ldo] bf16.
 * cin % 64 == 0, n_pad % 16 == 0 (16..256), ldo % 16 == 0, taps <= 9.  relu_out: ReLU on the result.
 * flags: 0 = default kernel (two-CTA tcgen05 pair when n_pad % 32 == 0, tap-major K); bit 0 is reserved (a ReLU on the
 * input is not supported: apply relu_out in the producing layer); bit 1 = taps innermost in K with L1-allocating gathers;
 * bit 2 = one CTA per 256-row tile (cta_group::1); bit 3 = deeper publish lag in the pair kernel (cp.async gathers);
 * bit 4 = no small-batch forms (128-row tiles with one accumulator, two producer teams; output channels split over two
 * CTAs); bit 5 = 16-byte cp.async gathers in the large-batch pair form instead of the TMA row gather
 * (cp.async.bulk.tensor tile::gather4).  The variants compute the same function and exist for comparison. */
int nz_hexconv_bf16(const void* x, const int32_t* nbr, const void* wt, const void* residual, void* out, int rows, int cells,
                    int taps, int cin, int n_pad, int ldo, int flags, int relu_out, void* stream);

/* Profiling aid: device buffer (int64[4 * chunks + 4]) that CTA 0 of the next nz_hexconv_bf16 launches fills with
 * clock64 stamps (stage free, chunk published, MMA start per K chunk; epilogue start/end).  NULL switches it off. */
int nz_hexconv_set_trace(void* dev_buffer);

/* n draws of the device root-noise generator (Philox4x32-10 + Marsaglia-Tsang Gamma(alpha, scale)), the
 * throughput-mode replacement of np.random.gamma in Explorer.add_exploration_noise (Explorer.py:208). */
int nz_noise_probe(double* out, int n, double alpha, double scale, uint64_t seed, void* stream);

/* words per slot in the "ctl" buffer and their meaning.  Words 0-15 are the block the search kernel loads at its start
 * (two 256-bit loads); words 0-7 are the ones it writes back (one 256-bit store). */
#define NZ_CTL_WORDS 32
enum {
  NZ_CTL_PHASE = 0, NZ_CTL_POOL_TOP = 1, NZ_CTL_SIMS_DONE = 2,
  NZ_CTL_ROOT_N0 = 3,   /* visit count the root had when it became the root (root N = ROOT_N0 + SIMS_DONE [+ pending leaves]) */
  NZ_CTL_ROOT_K = 4,    /* number of children of the root (they live at nodes 1 .. ROOT_K of the slot's pool) */
  NZ_CTL_PATH_LEN = 5, NZ_CTL_LEAF = 6, NZ_CTL_ERROR = 7,
  NZ_CTL_NOISED = 8, NZ_CTL_MAP = 9, NZ_CTL_MOVE = 10, NZ_CTL_UID = 11, NZ_CTL_GAMES_DONE = 12, NZ_CTL_CHOSEN = 13,
  NZ_CTL_HALF = 14,     /* compaction: which half of the general pool holds the current tree */
  NZ_CTL_N_PENDING = 15,
  NZ_CTL_ROOT = 16,     /* node index of the root: always 0 (kept for bindings that read it) */
  /* running totals for roofline accounting (u32 wraps are handled by the host reading deltas) */
  NZ_CTL_N_SIMS = 17, NZ_CTL_N_LEVELS = 18, NZ_CTL_N_SCANNED = 19, NZ_CTL_N_EXPAND = 20,
  NZ_CTL_N_CREATED = 21, NZ_CTL_N_MOVES = 22, NZ_CTL_N_TERMINAL = 23,
  NZ_CTL_LEAF_ROW = 24,     /* dense rows: the row of the leaf tensor the slot's pending leaf was written to */
  NZ_CTL_N_CACHE_HITS = 25, /* leaves expanded from the in-kernel inference cache */
  NZ_CTL_N_CACHE_SHARED = 26, /* leaves that waited for another game's network row of the same launch (same state) */
  NZ_CTL_LEAF_SLOT = 27      /* dense rows: table slot + 1 of the cache entry the pending leaf belongs to (0: none) */
};

#ifdef __cplusplus
}
#endif
#endif /* NZ_ENGINE_H */
