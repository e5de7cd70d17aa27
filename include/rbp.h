/* rbp.h — C ABI of librbp_b200.so: the B200-native (sm_100a) drop-in for robopoker's data-parallel
 * training path.  Every entry point cites the reference seam it replaces (paths relative to the
 * krukah/robopoker checkout).  Plain pointers and sizes only; the library owns all device state
 * behind opaque handles; every call is synchronous with respect to the HOST buffers it is given.
 * All functions return RBP_OK (0) or a negative status; the reference's convention on this path is
 * to panic (`expect`) — the Rust shim in INTEGRATION.md panics on any non-zero status.
 * There is no CPU fallback: every compute entry point fails with RBP_ERR_NO_DEVICE without a GPU.
 */
#ifndef RBP_H
#define RBP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    RBP_OK = 0,
    RBP_ERR_INVALID = -1,    /* bad argument / unknown enum */
    RBP_ERR_NO_DEVICE = -2,  /* no CUDA device (the library never computes on the CPU) */
    RBP_ERR_CUDA = -3,       /* a CUDA call failed; see rbp_last_error() */
    RBP_ERR_CAPACITY = -4,   /* game / buffer exceeds a compiled-in or caller-provided capacity */
    RBP_ERR_STATE = -5       /* call sequence error (e.g. step before init) */
};
const char* rbp_status_string(int status);
const char* rbp_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches claim) */
uint64_t rbp_kernel_launches(void);
int rbp_device_count(void);

/* ─────────────────────────────── multi-GPU: one process per GPU ─────────────────────────────── */

/* The reference's fast path is one rayon process (crates/mccfr/src/solver/solver.rs:225-240, crates/elkan/src/elkan.rs:80-168);
 * SURVEY §8b/§8e shard it across the GPUs of one box.  A communicator is this process's rank in that job.  The exchange
 * step of every sharded path runs INSIDE the library on the handle's stream once a communicator is attached
 * (rbp_nlhe_attach_comm, rbp_kmeans_attach_comm, rbp_solver_attach_comm): the host only distributes the 128-byte id.
 * NCCL (libnccl.so.2, located with dlopen: RBP_NCCL_LIB, then the loader path) carries the bootstrap, the barriers and
 * the plain collectives; the NLHE record/row exchange is written by the library's own kernels into peer memory mapped
 * over NVLink (CUDA IPC).  rbp_comm_unique_id is called on ONE rank; every rank then calls rbp_comm_init with that id. */
typedef struct rbp_comm rbp_comm_t;
int rbp_comm_unique_id(uint8_t out[128]);
int rbp_comm_init(int world_rank, int world_size, const uint8_t id[128], int device, rbp_comm_t** out);
void rbp_comm_destroy(rbp_comm_t* c); /* after every handle attached to it */
int rbp_comm_rank(rbp_comm_t* c);
int rbp_comm_size(rbp_comm_t* c);
int rbp_comm_barrier(rbp_comm_t* c); /* host-blocking: all ranks have reached it */

/* ─────────────────────────────── MCCFR ─────────────────────────────── */

/* crates/mccfr/src/solver/encounter.rs:21-27 — one (infoset, action) row, 16 B, same field order */
typedef struct {
    float weight;
    float regret;
    float payoff;
    uint32_t visits;
} rbp_encounter_t;

/* Row as exchanged with the host: key = (packed infoset key, action index in `choices()` order).
 * Packed infoset keys (also a Philox counter word of the RNG contract below):
 *   Kuhn  (crates/kuhn/src/info.rs:9-76):   bit0 acting | bits1-2 History{Open,Check,Bet,CheckBet} | bits3-4 Rank{J,Q,K}
 *   Leduc (crates/leduc/src/info.rs:12-94): bit0 acting | bits1-2 board (0 none, 1+Rank) | bits3-4 r1 Spot
 *                                           | bits5-7 r2 (0 none, 1+Spot) | bits8-9 own Rank
 *   Spot = {Open,Checked,Raised,CheckRaised} (crates/leduc/src/game.rs:6-11)                            */
typedef struct {
    uint32_t info_key;
    uint32_t action;
    rbp_encounter_t row;
} rbp_profile_row_t;

/* generic choices R, W, S of `trait Solver` (crates/mccfr/src/solver/solver.rs:38-74) as enums */
enum { RBP_GAME_KUHN = 0, RBP_GAME_LEDUC = 1, RBP_GAME_RPS = 2 /* crates/roshambo: info_key = bit0 acting | bits1-2 player */ };
enum { RBP_REGRET_SUMMED = 0, RBP_REGRET_FLOORED = 1, RBP_REGRET_LINEAR = 2, RBP_REGRET_DISCOUNTED = 3, RBP_REGRET_ASYMMETRIC = 4 }; /* crates/mccfr/src/regret/ (every schedule file) */
enum { RBP_WEIGHT_CONSTANT = 0, RBP_WEIGHT_LINEAR = 1, RBP_WEIGHT_QUADRATIC = 2, RBP_WEIGHT_EXPONENTIAL = 3 };                       /* crates/mccfr/src/policy/ (every schedule file) */
enum { RBP_SAMPLING_EXTERNAL = 0, RBP_SAMPLING_VANILLA = 1, RBP_SAMPLING_PRUNABLE = 2, RBP_SAMPLING_PLURIBUS = 3, RBP_SAMPLING_TARGETED = 4 };                  /* crates/mccfr/src/sample/ (every scheme file) */
/* how the per-tree Decisions of one epoch are folded into the table:
 *   ORDERED — reference semantics (solver.rs:96-105): one schedule application per Decisions, in tree order.
 *   BATCHED — one schedule application per row per epoch on the blocked-order sum of the deltas
 *             (the "allreduce of deltas" form; identical to ORDERED at batch 1).                       */
enum { RBP_FOLD_ORDERED = 0, RBP_FOLD_BATCHED = 1 };

/* process-global hyper-parameter singletons of the reference, as one POD:
 * crates/mccfr/src/hyperparams/{sampling.rs:39-50, pruning.rs:40-55, training.rs:52-60} */
typedef struct {
    float temperature;     /* 1.0  */
    float smoothing;       /* 2.0  */
    float curiosity;       /* 0.05 */
    float prune_threshold; /* -3e5 */
    float prune_explore;   /* 0.05 */
    uint32_t prune_warmup; /* 16384 */
    float regret_min;      /* -4e6 */
} rbp_hyper_t;
void rbp_hyper_default(rbp_hyper_t* out);

/* RNG contract (replaces SipHash+SmallRng of crates/mccfr/src/strategy/flow.rs:285-295 and the thread RNG
 * of `CfrGame::root`): Philox4x32-10, key=(seed_lo,seed_hi), counter=(epoch, tree_id, info_key, tag),
 * tag 0 node draw / 1 root deal (info_key 0xFFFFFFFF) / 2 pluribus coin;
 * range(n)=(u64(r0)*n)>>32; unit=(r0>>8)*2^-24; weighted = first i with unit*Σw < Σ_{j<=i}w_j (sequential f32). */
void rbp_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

typedef struct rbp_solver rbp_solver_t;

/* `Kuhn::<R,W,S>::default()` / `Leduc::<R,W,S>::default()` (crates/mccfr/src/strategy/macros.rs:7-151) with
 * `batch_size()` = batch.  `world_rank/world_size` shard the trees of an epoch: this handle samples tree ids
 * [rank*batch, (rank+1)*batch) of a global batch of world_size*batch trees (see rbp_solver_* exchange calls). */
int rbp_solver_create(int game, int regret, int weight, int sampling, int fold_mode, int batch, uint64_t seed,
                      const rbp_hyper_t* hyper /* NULL = defaults */, int device, rbp_solver_t** out);
int rbp_solver_set_world(rbp_solver_t* s, int world_rank, int world_size);
/* run this handle's kernels on a caller-owned cudaStream_t (e.g. the stream a collective library uses) */
int rbp_solver_set_stream(rbp_solver_t* s, void* cuda_stream);
void rbp_solver_destroy(rbp_solver_t* s);
/* `Solver::step` ×n (solver.rs:96-105) — sample `batch` trees, compute Decisions, fold, advance epoch */
int rbp_solver_step(rbp_solver_t* s, uint64_t n_epochs);
/* `Solver::spend` (solver.rs:130-137): step in a tight loop until `seconds` of wall clock are used; returns the epochs run and the time taken
 * (the per-decision refinement budget of the real-time players, crates/subgame) */
int rbp_solver_spend(rbp_solver_t* s, double seconds, uint64_t* epochs_out, double* elapsed_out);
/* rbp_solver_step with CUDA-event timing on the library's own stream (bench.py): every epoch is bracketed by
 * events, optionally preceded by an (untimed) L2 flush; returns the summed device milliseconds of the n steps and,
 * if non-NULL, the summed durations of the sampling and fold kernels. */
int rbp_solver_step_timed(rbp_solver_t* s, uint64_t n_epochs, int flush_l2, float* ms_total, float* ms_sample, float* ms_fold);
/* device self-test: the fold kernel's reciprocal-based Welford division (payoff += (x - payoff)/(visits+1),
 * solver.rs:174-181) against IEEE division for every count in [1, max_count] x `samples` dividends */
int rbp_selftest_div_by_count(uint32_t max_count, uint32_t samples, uint64_t* mismatches);
/* `RefProf::t` (crates/mccfr/src/strategy/profile.rs:14) */
int rbp_solver_epochs(rbp_solver_t* s, uint64_t* out);
/* `Solver::exploitability` (solver.rs:327-338 → crates/mccfr/src/strategy/nash.rs:31-193) */
int rbp_solver_exploitability(rbp_solver_t* s, float* out);
/* telemetry of crates/mccfr/src/metrics/mod.rs: [0] nodes, [1] infosets (Decisions), [2] infoset-action regret updates */
int rbp_solver_counters(rbp_solver_t* s, uint64_t out[3]);
/* bulk forms of `RefProf::cum_*` / `MutProf::mut_*` / `CfrData::encounters_{ref,mut}`
 * (crates/mccfr/src/strategy/{profile.rs:12-25,storage.rs,book.rs:14-24}).  Export returns the rows the
 * reference's HashMap would hold (touched rows), sorted by (info_key, action). */
int rbp_profile_export(rbp_solver_t* s, rbp_profile_row_t* rows, int cap, int* n_out);
int rbp_profile_import(rbp_solver_t* s, const rbp_profile_row_t* rows, int n, uint64_t epochs);
/* `RefProf::averaged_distribution` (profile.rs:41-45) for one infoset; returns n actions in *n_out */
int rbp_profile_averaged(rbp_solver_t* s, uint32_t info_key, float* probs, int cap, int* n_out);
/* shape of the enumerated game: [0] nodes, [1] terminals, [2] decision infosets, [3] rows,
 * [4] max nodes of a sampled tree, [5] max walker infosets per sampled tree */
int rbp_solver_game_shape(rbp_solver_t* s, int out[6]);

/* ---- safe subgame solving on the small games: `WorldSolver` (crates/subgame/src/world/solver.rs:33-146), which is also
 * `SubGameSolver` without an origin (crates/subgame/src/solver.rs:46-146; depth-limited frontiers are NOT built) ----
 * RNG contract (the reference samples the world from the unseeded thread RNG): world = weighted(belief weights) on
 * Philox counter (step, 0, 0xFFFFFFFE, tag 5); the step's tree draws with tree_id = world. */
typedef struct rbp_subgame rbp_subgame_t;
/* `Partition::partition::<W>` (world/partition.rs:27-53): posterior reach of every secret (ascending secret order) -> its world
 * (0 = highest reach) and the probability mass of every world */
int rbp_subgame_partition(const float* reach, int n_secrets, int worlds, int32_t* world_of_secret, float* weights);
/* the action-conditioned posterior over the external player's rank (kuhn/src/solver.rs `subgame_with_reach_conditioned_posterior`): per card
 * the external player could hold, `Solver::external_reach` (mccfr/src/solver/solver.rs:198-211) along the path under the blueprint's averaged
 * policy, summed per rank (`Posterior::add`).  `rows` = rbp_profile_export of the blueprint.  Host arithmetic, no device needed. */
int rbp_subgame_posterior(int game, const rbp_profile_row_t* rows, int n_rows, int external, int c0, int c1, const uint8_t* path, int path_len,
                          float* reach3);
/* `WorldSolver::new(encoder, profile, external, belief, recall)`: `blueprint` = a trained Kuhn / Leduc solver (its table is copied: the
 * blueprint may be destroyed afterwards); belief = world of every rank (NULL = `Belief` without members: every secret is remembered)
 * + `worlds` weights; recall = the observed deal (c0, c1: card = 2 * rank + suit, `Card::ALL` order) and the path of branch indices
 * (`branches()` order; at a chance node: index among the cards still in the deck) from the dealt root to the entry state.
 * The tree of a step stops at chance nodes (world/encoder.rs:97-106); such a leaf is worth `frontier_payoff` of its nearest decision
 * ancestor (mccfr/src/strategy/nash.rs:50-79: the stored V(I), local or the blueprint's). */
int rbp_subgame_create(rbp_solver_t* blueprint, int external, int worlds, const int32_t* world_of_rank /* [3] or NULL */, const float* weights,
                       int c0, int c1, const uint8_t* path, int path_len, uint64_t seed, rbp_subgame_t** out);
void rbp_subgame_destroy(rbp_subgame_t* g);
/* host half of the constructor (no device needed): per world the restricted deal [8][2], the flat node and the infoset key of the entry state */
int rbp_subgame_entries(int game, int external, int worlds, const int32_t* world_of_rank, int c0, int c1, const uint8_t* path, int path_len,
                        int32_t* cards16, int32_t* nodes8, uint32_t* info_keys8);
/* `WorldSolver::step` x n (world/solver.rs:118-146): sample a world, `WorldRestrict::restrict` (kuhn/src/encoder.rs:47-66,
 * leduc/src/encoder.rs:48-70), one ExternalSampling tree, SummedRegret + LinearWeight fold into the world's table */
int rbp_subgame_step(rbp_subgame_t* g, uint64_t n);
/* `Solver::spend` (mccfr/src/solver/solver.rs:130-137) */
int rbp_subgame_spend(rbp_subgame_t* g, double seconds, uint64_t* steps_out, double* elapsed_out);
/* steps run (`WorldProfile::t`), how often each world was drawn [8], the restricted entry deal of each world [8][2] */
int rbp_subgame_info(rbp_subgame_t* g, uint64_t* steps, uint64_t* drawn8, int32_t* entry_cards16);
/* `WorldProfile::local` of one world (world/profile.rs:25-36): the rows the subgame wrote, sorted by (info_key, action) */
int rbp_subgame_export(rbp_subgame_t* g, int world, rbp_profile_row_t* rows, int cap, int* n_out);
/* `CfrNash::averaged_policy` over `WorldInfo(world, info)`: local weights, the blueprint's where the edge was never written */
int rbp_subgame_averaged(rbp_subgame_t* g, int world, uint32_t info_key, float* probs, int cap, int* n_out);
/* `Harvest::harvest(base)` (world/solver.rs:148-191): refined policy (mean over the worlds of the regret-matching policy), visits
 * summed over the worlds, positive regret summed over edges and worlds */
int rbp_subgame_harvest(rbp_subgame_t* g, uint32_t info_key, float* refined, uint32_t* visits, float* regret, int cap, int* n_out);

/* Multi-GPU exchange (one process per GPU).  The library does not link a collective library: the host
 * (Rust shim / torch.distributed) moves these device buffers with NCCL.  BATCHED fold: after
 * rbp_solver_sample() each rank holds its blocked partial sums; all-gather `delta` buffers across ranks into
 * `gathered` (rank-major) and call rbp_solver_fold_gathered(), which sums in rank order — bit-identical on
 * every rank and to the single-GPU run with world_size*batch trees. */
int rbp_solver_attach_comm(rbp_solver_t* s, rbp_comm_t* c); /* the same exchange inside the library: rbp_solver_step then all-gathers on its stream */
int rbp_solver_sample(rbp_solver_t* s);                                   /* K1 only (+ local blocked sums)     */
int rbp_solver_delta_buffer(rbp_solver_t* s, void** dev_ptr, size_t* bytes); /* this rank's partial sums (device) */
int rbp_solver_fold_gathered(rbp_solver_t* s, const void* dev_gathered, int world_size); /* K2 over all ranks  */

/* ─────────────────────────────── MCCFR: heads-up no-limit hold'em blueprint ─────────────────────────────── */

/* `Nlhe<R,W,S>` (crates/nlhe/src/solver.rs:11 `mccfr!(Nlhe, NlheEncoder, NlheTurn, NlheEdge, NlheGame, NlheInfo, 128)`;
 * `Flagship` = LinearRegret + LinearWeight + PluribusSampling, crates/nlhe/src/lib.rs:86-90) with `batch_size()` = batch.
 * The game plug-in is compiled in: `kicker::Game` heads-up (STACK 200, blinds 1/2, crates/kicker/src/game.rs),
 * the Pluribus action grid (crates/pokerkit/src/lib.rs:60-160), `NlheGame::apply` with snapping
 * (crates/nlhe/src/game.rs:35-55), `NlheInfo` = (current-street subgame Path, choices Path, Abstraction)
 * (crates/nlhe/src/info.rs:141-160).  The profile is a device-resident open-addressing table of
 * `table_slots` (power of two, 0 = 2^22) infosets x up to 10 edges (replaces HashMap<NlheInfo, HashMap<NlheEdge, Encounter>>).
 * `max_nodes_per_tree`: node capacity of one epoch = batch x this (0 = automatic: 768 x batch + 16384; sampled trees
 * average ~450 nodes, the largest seen ~3500); an epoch that exceeds a capacity fails with RBP_ERR_CAPACITY.
 * RNG contract additions to rbp_philox4x32_10's: hole cards = Philox(epoch, tree, 0xFFFFFFFF, tag 1) words 0..3
 * through Deck::draw (crates/deuce/src/deck.rs:28-43, its bias kept); board cards = Philox(epoch, tree,
 * lo32(hist), tag 4), hist = running mix64 hash of the edges applied since the root; node draws use
 * info word = lo32(mix64(subgame ^ mix64(choices ^ mix64(abstraction)))).
 * Abstraction lookup (`NlheEncoder::abstraction`, crates/nlhe/src/encoder.rs:30-35): synthetic until a table is
 * installed — bucket = mix64(canonical pocket * 0x9E3779B97F4A7C15 ^ mix64(canonical public)) mod {169,256,256,101},
 * Abstraction = street << 8 | bucket (crates/kicker/src/abstraction.rs:15-60). */
typedef struct rbp_nlhe rbp_nlhe_t;
struct rbp_isoset; /* rbp_isoset_t, declared with the isomorphism entry points below */
int rbp_nlhe_create(int regret, int weight, int sampling, int batch, uint64_t seed, const rbp_hyper_t* hyper /* NULL = defaults */,
                    uint64_t table_slots, int max_nodes_per_tree, int device, rbp_nlhe_t** out);
void rbp_nlhe_destroy(rbp_nlhe_t* s);
int rbp_nlhe_set_world(rbp_nlhe_t* s, int world_rank, int world_size);
int rbp_nlhe_set_stream(rbp_nlhe_t* s, void* cuda_stream);
/* `Solver::step` x n (crates/mccfr/src/solver/solver.rs:96-105): sample `batch` trees, Decisions per walker infoset,
 * fold in tree order (one schedule application per Decisions), advance the epoch */
int rbp_nlhe_step(rbp_nlhe_t* s, uint64_t n_epochs);
/* `Solver::spend` for the NLHE solver (one epoch at a time) */
int rbp_nlhe_spend(rbp_nlhe_t* s, double seconds, uint64_t* epochs_out, double* elapsed_out);
/* rbp_nlhe_step with CUDA-event timing per phase (summed over the epochs): ms[0] total, [1] tree build (level expansion,
 * size/preorder sweeps, scatter), [2] value kernels, [3] resolve + radix sort, [4] fold, and with a communicator [5] records
 * to their owners (incl. barrier), [6] touched rows to every peer (incl. barrier) + install; [7] reserved.  flush_l2: one
 * untimed 192 MiB write before the first epoch of the call (with a communicator followed by a barrier, so ranks start together). */
int rbp_nlhe_step_timed(rbp_nlhe_t* s, uint64_t n_epochs, int flush_l2, float ms[8]);
/* Sharded `Solver::step`: rank r samples tree ids [r*batch, (r+1)*batch) of every epoch (global batch = world*batch trees,
 * folded in tree order exactly as one process would: solver.rs:96-105) and owns the infosets with hash(key) mod world == r.
 * Per epoch, on the handle's stream and without host involvement: update records are stored into their owner's memory by
 * the partition kernel itself, the owner folds them, the touched rows are stored into every peer's memory and installed —
 * after which every rank's table holds the rows of ONE process running world*batch trees, bit for bit.  Collective call.
 * After it, rbp_nlhe_step / rbp_nlhe_step_timed run the sharded epoch.  Needs table_slots <= 2^27. */
int rbp_nlhe_attach_comm(rbp_nlhe_t* s, rbp_comm_t* c);
/* out[0] epochs, [1] nodes, [2] Decisions ("infos"), [3] infoset-action regret updates, [4] table rows in use,
 * [5] update records of the last epoch, [6] largest tree of the run (nodes), [7] reserved */
int rbp_nlhe_counters(rbp_nlhe_t* s, uint64_t out[8]);
/* decision nodes visited while sampling, cumulative — the profile reads of SURVEY §8d's algorithmic-bytes model: out[0] walker
 * decision nodes, [1] sum of their choice counts A, [2] opponent decision nodes, [3] sum of their A */
int rbp_nlhe_traffic_counters(rbp_nlhe_t* s, uint64_t out[4]);
/* One row of the reference's blueprint table (crates/nlhe/src/profile.rs:143-160: past, present, choices, edge, weight,
 * regret, payoff, visits) */
typedef struct {
    int64_t past;    /* i64::from(info.subgame()) — Path, 5 bits per edge, first edge lowest (crates/kicker/src/path.rs) */
    int64_t choices; /* i64::from(info.choices()) */
    int64_t edge;    /* u64::from(Edge) as i64 (crates/kicker/src/edge.rs:185-197): Draw 0, Fold 1, Check 2, Call 3, Raise 4 | numer << 3 | denom << 11,
                      * Shove 5, Open 6 | n << 3.  Import also reads the legacy BBs form (tag 4, bit 19; edge.rs:168-172).  NOT the 5-bit path code. */
    int16_t present; /* i16::from(info.bucket()) — Abstraction */
    int16_t pad[3];
    rbp_encounter_t row;
} rbp_nlhe_row_t;
/* rows sorted by (past, present, choices, position of edge in choices); cap = capacity of `rows` (NULL to count) */
int rbp_nlhe_export(rbp_nlhe_t* s, rbp_nlhe_row_t* rows, uint64_t cap, uint64_t* n_rows);
int rbp_nlhe_import(rbp_nlhe_t* s, const rbp_nlhe_row_t* rows, uint64_t n_rows, uint64_t epochs);
/* `NlheEncoder(BTreeMap<Isomorphism, Abstraction>)` (crates/nlhe/src/encoder.rs:23-35; the table `Hydrate` streams from
 * Postgres, encoder.rs:187-214): install the abstraction lookup of ONE street from a device-resident isomorphism set
 * whose abstraction column is filled (rbp_isoset_river_buckets, or rbp_isoset_set_abstractions with the k-means
 * assignments of that street).  Streets without a table keep the synthetic lookup.  After this call an observation that
 * is missing from the table fails the step with RBP_ERR_STATE — the reference panics ("isomorphism not found"). */
int rbp_nlhe_set_lookup(rbp_nlhe_t* s, struct rbp_isoset* isos);
/* the same from host rows in the reference's `isomorphism` table format (obs = i64::from(Isomorphism), abs =
 * i16::from(Abstraction)); one call per street with every row of that street */
int rbp_nlhe_set_lookup_rows(rbp_nlhe_t* s, const int64_t* obs, const int16_t* abs, int64_t n);
/* multi-GPU exchange (one process per GPU; the library does not link a collective library): after rbp_nlhe_sample
 * this rank's update records sit in a device buffer (`words` 32-bit words per record, `count` of them); the host
 * all-gathers them and hands the concatenation (any rank order: the fold sorts by (infoset, tree)) to
 * rbp_nlhe_fold_records, which advances the epoch. */
int rbp_nlhe_sample(rbp_nlhe_t* s);
int rbp_nlhe_records(rbp_nlhe_t* s, void** device_ptr, uint64_t* count, uint64_t* capacity, int* words_per_record);
int rbp_nlhe_fold_records(rbp_nlhe_t* s, const void* device_records, uint64_t count);
/* Owner-sharded fold (scales the fold with the world): infosets are owned by rank hash(key) mod world.
 *   rbp_nlhe_sample → rbp_nlhe_partition_records (this rank's records grouped by owner; counts[r] go to rank r)
 *   → host all-to-all → rbp_nlhe_fold_records (the records this rank owns) → rbp_nlhe_touched_rows (key + 10 encounters per
 *   touched infoset, `words` 32-bit words each) → host all-gather → rbp_nlhe_apply_rows (overwrite the replica).
 * Every rank's table then holds the same rows as one process folding the whole epoch. */
int rbp_nlhe_partition_records(rbp_nlhe_t* s, void** device_ptr, uint64_t* counts /* [world_size] */);
int rbp_nlhe_touched_rows(rbp_nlhe_t* s, void** device_ptr, uint64_t* count, int* words_per_row);
int rbp_nlhe_apply_rows(rbp_nlhe_t* s, const void* device_rows, uint64_t count);
/* test hook: tree `tree` of the CURRENT epoch as the sampler builds it — per node (preorder, children in choices order)
 * depth, kind (0 walker 1 opponent 2 chance 3 terminal), action index, policy p, sampling q, terminal payoff */
typedef struct {
    uint8_t depth, kind, act, pad;
    float p, q, payoff;
} rbp_nlhe_node_t;
int rbp_nlhe_debug_tree(rbp_nlhe_t* s, int tree, rbp_nlhe_node_t* out, int cap, int* n_nodes);

/* ─────────────────────────────── deuce: hand strength and river equity ─────────────────────────────── */

/* `Strength::from(Hand)` (crates/deuce/src/strength.rs:19-31 → evaluator.rs:39-177) for a batch of 52-bit hands
 * (bit = 4*rank + suit, crates/deuce/src/{hand.rs,card.rs}; 5 to 7 cards).  Packed so that unsigned integer order
 * equals the reference's derived `Ord` on (Ranking, Kickers) (ranking.rs:33-44, default build: FullHouse < Flush):
 *   bits 24-27 tag {HighCard 0, OnePair 1, TwoPair 2, ThreeOAK 3, Straight 4, FullHouse 5, Flush 6, FourOAK 7,
 *   StraightFlush 8} | 20-23 first rank | 16-19 second rank | 0-12 `Kickers` rank bits (kicks.rs:4).          */
int rbp_eval_batch(const uint64_t* hands, int64_t n, uint32_t* strength_out);
/* `Observation::equity` (crates/deuce/src/observation.rs:45-62) for n river observations (pocket: 2 cards, public:
 * 5 cards): wins/(wins+losses) over all C(45,2)=990 villain holes, ties dropped, 0.5 if nothing decisive; bucket =
 * `Abstraction::from(equity)` index = round(100 p) (crates/kicker/src/abstraction.rs:43-45,61-63).  Any output
 * pointer may be NULL.  This is the per-isomorphism work of `Lookup::grow(Street::Rive)` (lloyd/src/lookup.rs:177-184). */
int rbp_river_equity_batch(const uint64_t* pocket, const uint64_t* pub, int64_t n, float* equity_out, uint8_t* bucket_out,
                           uint32_t* wins_out, uint32_t* total_out);
/* same, on device buffers and a caller stream (cudaStream_t passed as void*) — no copies, asynchronous */
int rbp_river_equity_device(const uint64_t* d_pocket, const uint64_t* d_public, int64_t n, float* d_equity, uint8_t* d_bucket,
                            uint32_t* d_wins, uint32_t* d_total, void* stream);

/* ─────────────────────────────── deuce: suit isomorphisms, lookups, projections ─────────────────────────────── */

typedef struct rbp_isoset rbp_isoset_t;
/* `IsomorphismIterator::from(street)` (crates/deuce/src/isomorphism_iter.rs:7-21): the canonical observations of a
 * street (0 Pref, 1 Flop, 2 Turn, 3 Rive) in the reference's enumeration order, device-resident, sorted by
 * (pocket, public).  Sizes: 169 / 1,286,792 / 13,960,050 / 123,156,254 (crates/deuce/src/street.rs:129-135). */
int rbp_isoset_create(int street, int device, rbp_isoset_t** out);
void rbp_isoset_destroy(rbp_isoset_t* h);
int64_t rbp_isoset_size(rbp_isoset_t* h);
/* copy out a slice: 52-bit card sets (`u64::from(Hand)`), and the abstraction column if attached (nullable pointers) */
int rbp_isoset_export(rbp_isoset_t* h, int64_t offset, int64_t count, uint64_t* pocket_out, uint64_t* public_out, uint8_t* abs_out);
/* attach the `Lookup` column isomorphism → abstraction index (e.g. `rbp_kmeans_assign` output narrowed to u8) */
int rbp_isoset_set_abstractions(rbp_isoset_t* h, const uint8_t* abs);
/* `i64::from(Observation)` and back (crates/deuce/src/observation.rs:130-163; host-side format conversion): board cards then
 * pocket cards in ascending card order, one byte (1 + card) each, first card in the highest used byte */
void rbp_obs_encode(const uint64_t* pocket, const uint64_t* public_, int64_t n, int64_t* obs_out);
void rbp_obs_decode(const int64_t* obs, int64_t n, uint64_t* pocket_out, uint64_t* public_out);
/* the (obs i64, abs i16) rows the reference stores in its `isomorphism` table (crates/lloyd/src/lookup.rs `Streamable::rows`) */
int rbp_isoset_export_rows(rbp_isoset_t* h, int64_t offset, int64_t count, int64_t* obs_out, int16_t* abs_out);
/* `Lookup::grow(Street::Rive)` (crates/lloyd/src/lookup.rs:177-184): column = `Abstraction::from(equity)` of every river iso */
int rbp_isoset_river_buckets(rbp_isoset_t* h);
/* `Lookup::projections` (lookup.rs:46-66): for parent observations [offset, offset+count) the histogram over the child
 * street's abstractions of their children (`Observation::children` → `Isomorphism::from` → lookup → `Histogram::increment`):
 * hist_out[count][bins] u8.  *misses_out counts children not found in the child set (0 when the set is complete). */
int rbp_isoset_project(rbp_isoset_t* parent, rbp_isoset_t* child, int bins, int64_t offset, int64_t count, uint8_t* hist_out, uint64_t* misses_out);
/* `Isomorphism::from(Observation)` (crates/deuce/src/isomorphism.rs:9-15, permutation.rs:9-66) for a batch;
 * flag_out (nullable) = `Isomorphism::is_canonical` */
int rbp_canonical_batch(const uint64_t* pocket, const uint64_t* pub, int64_t n, uint64_t* pocket_out, uint64_t* public_out, uint8_t* flag_out);

/* ─────────────────────────────── lloyd / elkan: k-means abstraction layers ─────────────────────────────── */

/* distance kinds of `Metric::emd` (crates/lloyd/src/metric.rs:109-115) */
enum {
    RBP_KMEANS_W1 = 0,      /* Equity::variation over the 101 river-equity buckets (turn layer) */
    RBP_KMEANS_SINKHORN = 1 /* Sinkhorn::divergence over next-street clusters with a ground metric (flop layer) */
};

typedef struct rbp_kmeans rbp_kmeans_t;

/* `Layer::<K,N>::build` (crates/lloyd/src/layer.rs:250-272): N points = dense histograms `counts[n][bins]` (u8 counts;
 * the reference's `Bins{weight, counts:[usize;N]}`, lloyd/src/bins.rs:29-38 — weight = Σ counts), K clusters.
 * The points are copied to the device once and stay resident. */
int rbp_kmeans_create(int kind, int64_t n, int k, int bins, const uint8_t* counts, int device, rbp_kmeans_t** out);
void rbp_kmeans_destroy(rbp_kmeans_t* h);
/* SINKHORN layers only: the ground `Metric` of the next street (crates/lloyd/src/metric.rs:25-31; `Layer::build`
 * loads it with `Metric::from_street`, layer.rs:262): tri[bins(bins-1)/2] in `Pair::merge` order, diagonal implied 0.
 * Must be set before centroids; also computes the memoised self terms OT(x,x) of every point (sinkhorn.rs:172-191). */
int rbp_kmeans_set_metric(rbp_kmeans_t* h, const float* tri, int bins);
/* `Layer::init_centroids` k-means++ (layer.rs:140-181).  The reference's SmallRng/WeightedIndex<f32> stream is
 * replaced by the integer-weight contract: round r draws word = Philox(counter=(r,0,0,3), key=seed),
 * q_i = (u64)(min(potential_i, 2^20) * 2^32), T = Σ q_i, x = mulhi64(word, T), pick = first i with x < Σ_{k<=i} q_k.
 * chosen_out[k] (nullable) receives the chosen point indices. */
int rbp_kmeans_init_pp(rbp_kmeans_t* h, uint64_t seed, int32_t* chosen_out);
/* explicit centroids: `counts[k][bins]` (u64 sums of member counts, as `Absorb` produces them, elkan/src/absorb.rs:16-21) */
int rbp_kmeans_set_centroids(rbp_kmeans_t* h, const uint64_t* counts);
/* `Elkan::init_bounds` (crates/elkan/src/elkan.rs:39-47): argmin_j distance(c_j, x) with first-minimum ties */
int rbp_kmeans_init_bounds(rbp_kmeans_t* h);
/* `Elkan::step_elkan` (elkan.rs:153-168) + `Prior::tally` (elkan/src/prior.rs:35-47): drift_out[k], sizes_out[k],
 * *reassigned_out are nullable */
int rbp_kmeans_step(rbp_kmeans_t* h, float* drift_out, uint32_t* sizes_out, uint32_t* reassigned_out);
/* the same step split at the one exchange point for point-sharded multi-GPU runs: after _local each rank holds
 * integer partial sums (K x (bins+1) u64: counts + weight) plus sizes[k]/reassigned u32 counters on the device —
 * all-reduce(sum) them in place (integer: order-independent, bit-identical on every rank), then call _finish. */
int rbp_kmeans_step_local(rbp_kmeans_t* h);
int rbp_kmeans_accumulator(rbp_kmeans_t* h, void** dev_ptr, size_t* bytes);
int rbp_kmeans_counters(rbp_kmeans_t* h, void** dev_sizes /* u32[k] */, void** dev_reassigned /* u32[1] */);
void* rbp_kmeans_stream(rbp_kmeans_t* h); /* cudaStream_t the kernels run on */
int rbp_kmeans_step_finish(rbp_kmeans_t* h, float* drift_out, uint32_t* sizes_out, uint32_t* reassigned_out);
/* Point-sharded layers inside the library: after this call rbp_kmeans_step runs the point pass, ONE integer all-reduce (member sums, cluster
 * sizes and the reassignment counter in one buffer) on the layer's stream, and the centroid update — identical centroids on every
 * rank, bit for bit.  Every rank holds its own shard of the points and the same initial centroids (rbp_kmeans_set_centroids). */
int rbp_kmeans_attach_comm(rbp_kmeans_t* h, rbp_comm_t* c);
/* `Layer::lookup` (layer.rs:44-60): fresh naive argmin against the current centroids; dist_out nullable */
int rbp_kmeans_assign(rbp_kmeans_t* h, uint32_t* assign_out, float* dist_out);
/* `Layer::future` payload: centroid histograms (counts[k][bins], weights[k]); either may be NULL */
int rbp_kmeans_centroids(rbp_kmeans_t* h, uint64_t* counts_out, uint64_t* weights_out);
/* `Layer::metric` (layer.rs:85-101) + `Metric::from` normalisation (metric.rs:127-141): tri_out[k(k-1)/2] in the
 * triangular order of `Pair::merge` (lloyd/src/pair.rs:36-39): index(i>j) = i(i-1)/2 + j */
int rbp_kmeans_metric(rbp_kmeans_t* h, float* tri_out);
/* `Bounds<K>` state per point (elkan/src/bounds.rs:19-28) for inspection: assign[n], upper[n], lower[n][k], stale[n] */
int rbp_kmeans_bounds(rbp_kmeans_t* h, uint32_t* assign_out, float* upper_out, float* lower_out, uint8_t* stale_out);
/* measurement utility: sustained independent FP32 adds per second (in 1e12/s) on this GPU — the issue ceiling the
 * W1 distance kernels (202 FADD per distance, no FMA possible) are reported against */
int rbp_measure_fadd_peak(float* tera_adds_per_s);
/* CUDA-event timing on the library stream: what = 0 full step, 1 N x K assignment sweep */
int rbp_kmeans_timed(rbp_kmeans_t* h, int what, int iters, float* ms_out);
/* SINKHORN layers: work counters since creation / the last reset — out3 = {OT solves, Gauss-Seidel sweeps,
 * exp terms = Σ over solves of nx·ny·(2·sweeps + 1)} (the unit the flop layer's throughput is reported in); out3 nullable */
int rbp_kmeans_sinkhorn_stats(rbp_kmeans_t* h, uint64_t* out3, int reset);

/* SINKHORN layers: tensor-core screen of the two naive N x K sweeps — `Elkan::init_bounds` (crates/elkan/src/elkan.rs:39-47,68-77) and
 * `Layer::lookup` (crates/lloyd/src/layer.rs:62-84) — which need per point only argmin_j and that one distance.  With margin >= 0,
 * rbp_kmeans_init_bounds / rbp_kmeans_assign first compute an APPROXIMATE divergence of every point to every centroid in the
 * scaling domain (u = e^phi, v = e^psi, Gibbs kernel exp(-C/T); the softmin half-steps of sinkhorn.rs:96-129 as bf16 hi+lo x bf16
 * GEMMs on tcgen05 tensor cores with TMEM accumulators and a TMA-staged centroid tile: csrc/sk_screen.cuh) and then run the exact
 * log-domain solver only for the centroids within `margin` of the point's smallest approximate value, in centroid order with
 * the reference's first-minimum rule.  Assignments and winning distances equal the full exact sweep bit for bit whenever
 * margin >= 2 x the screen's error; rbp_kmeans_screen_probe measures that error (approximate divergences of points [0, m) to all
 * k centroids: out[m][k]; stats2 = {(point, 128-centroid tile) problems, Sinkhorn iterations summed}).  margin < 0 switches it off. */
int rbp_kmeans_screen(rbp_kmeans_t* h, float margin);
int rbp_kmeans_screen_probe(rbp_kmeans_t* h, int64_t m, float* out, uint64_t* stats2 /* nullable */);

/* `Metric::emd` / `Sinkhorn::divergence` (crates/lloyd/src/metric.rs:109-115, sinkhorn.rs:166-171) for explicit pairs:
 * out[t] = max(0, OT(A[ia[t]], B[ib[t]]) - OT(A,A)/2 - OT(B,B)/2) over dense u32 histograms a_counts[na][bins],
 * b_counts[nb][bins].  Log-domain Sinkhorn exactly as sinkhorn.rs:77-139 (defaults T=0.025, 128 iterations, tol 5e-4,
 * hyperparams/sinkhorn.rs:17-23), sequential sums in ascending-bucket order.  exp/ln contract (replaces the platform
 * libm of `f32::exp`/`f32::ln`, which no reference test pins): exp_c(x) with x clamped to [-87.33654, 88.72283]
 * (saturating; the softmin clamps at MIN_POSITIVE anyway): t = fma(x, 1.44269504, 1.5*2^23), k = t - 1.5*2^23,
 * r = fma(k, 2.12194440e-4, fma(k, -0.693359375, x)), P5 = Cephes expf coefficients in Horner form (one fma per
 * step), y = fma(P5, r*r, r) + 1, result bits = bits(y) + (bits(t) << 23); ln_c = Cephes logf on m in
 * [sqrt(1/2), sqrt 2), Horner and tail steps as fma.  Fixed IEEE operation sequences: nothing else is contracted. */
int rbp_sinkhorn_batch(const uint32_t* a_counts, int na, const uint32_t* b_counts, int nb, int bins, const int32_t* ia, const int32_t* ib,
                       int64_t n, const float* tri, float temperature, int iterations, float tolerance, float* out);

#ifdef __cplusplus
}
#endif
#endif /* RBP_H */
