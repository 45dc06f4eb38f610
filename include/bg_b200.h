/*
 * bg_b200.h -- C ABI of libbg_b200.so: the B200 (sm_100a) batched backgammon engine.
 *
 * Drop-in boundary for ONE hot path of Nick-qsv/MLP-PPO-2PLY-P3 (all-Python reference, no FFI of
 * its own): the functions below are what a ctypes/cffi binding on the reference side calls in place
 * of its per-game Python objects (INTEGRATION.md shows the stub).  Paths cite /root/reference/src.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch allocates), unless named host_*;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream), allocates nothing and never synchronises the host;
 *   - return value: BG_OK (0) or a negative BG_ERR_* code; bg_last_error() gives the message of the
 *     calling thread's last failure.  Data-dependent conditions (bad board, scratch/output overflow)
 *     cannot be returned synchronously: kernels OR them into the caller's device `status` word
 *     (BG_STATUS_*), which the caller reads when it next synchronises.  Nothing is dropped silently.
 *   - position layout "board52": int8[52] = [P1 points 0..23][P2 points 0..23][bar1 bar2 off1 off2],
 *     i.e. rows 0,1 and the used cells of rows 2,3 of the reference's (4,24) int8 tensor
 *     (board/immutable_board.py:20-27).  Counts must be 0..15.
 *   - players: int8, 0 = PLAYER1 (moves 0->23), 1 = PLAYER2 (moves 23->0)  (players/player.py:6-12).
 *   - dice: int8[2] per position, 1..6, any order (moves/get_all_moves.py:28-30 sorts internally).
 */
#ifndef BG_B200_H
#define BG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BG_OK 0
#define BG_ERR_INVALID (-1)   /* bad argument (null pointer, negative size, workspace too small) */
#define BG_ERR_CUDA (-2)      /* CUDA runtime error at launch; see bg_last_error() */

#define BG_STATUS_BAD_INPUT 1         /* a board had a count outside 0..15 or dice outside 1..6 */
#define BG_STATUS_SCRATCH_OVERFLOW 2  /* a position produced more boards per level than BG_MOVEGEN_CAP_BIG */
#define BG_STATUS_OUTPUT_OVERFLOW 4   /* the afterstate buffer was too small; affected counts are 0 */
#define BG_STATUS_DICE_EXHAUSTED 8    /* an external dice stream ran out */

#define BG_FEATURES 198               /* board/immutable_board.py:171-212 */
#define BG_BOARD_BYTES 52
#define BG_HIDDEN 128                 /* agent/config.py HIDDEN_SIZE, agent/policy_network.py:44 */
#define BG_FEAT_LD_BF16 208           /* 198 padded to a multiple of 16 (one tcgen05 K step) */
#define BG_ACTIONS 500                /* action slots = max_legal_moves (agent/train.py: action_size=500) */

const char* bg_last_error(void);
int bg_version(void);

/* ------------------------------------------------------------------------------------------------
 * K1  legal-move generation.  Replaces get_all_possible_moves (moves/get_all_moves.py:9-94) and
 * everything below it (moves/handle_moves.py:109-341, moves/move_logic.py:20-275,
 * moves/conditions.py:7-147, board/immutable_board.py:42-89,236-246) for B positions at once.
 * Afterstates of a position are emitted in the reference's legal_moves order, so index k here is
 * action k of the reference env (environment/backgammon_env.py:152).
 *
 * workspace: bg_movegen_workspace_bytes(B) bytes of device scratch (contents undefined).
 * counts_true[b] = number of legal plays (never truncated; -1 if the position was rejected).
 * row_features_bf16 (optional): K3 fused into K1's output stage -- the 208-wide bf16 feature row of every
 * afterstate written (turn flag = the mover, ai/batching.py:72-74), same row index as afterstates52.
 */
size_t bg_movegen_workspace_bytes(long long B);

/* pass 1 of the deterministic two-pass form: counts only */
int bg_movegen_count(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long B,
                     int32_t* counts_true, int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* pass 2: rows of position b are written at afterstates52 + 52*offsets[b] (offsets = exclusive scan
 * of min(counts_true, max_rows_per_board) computed by the caller; max_rows_per_board 0 = uncapped). */
int bg_movegen_write(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long B,
                     const long long* offsets, int max_rows_per_board, int8_t* afterstates52,
                     long long afterstate_capacity_rows, int8_t* row_players /*nullable: mover of each row*/,
                     uint16_t* row_features_bf16 /*nullable: fused K3, [rows][208] bf16*/, int32_t* counts_true /*nullable*/, int32_t* counts /*nullable*/, int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* single-pass form used by the env: every warp reserves its rows with one atomicAdd on *alloc_rows
 * (caller zeroes it -- or lets it live in bytes [48, 56) of `workspace`, which every K1 call zeroes itself together with its own
 * counters: one fill less per turn); starts[b] receives the first row.  Row blocks of different positions are in
 * arbitrary order, rows inside a block are in reference order.  counts[b] = min(true, max_rows). */
int bg_movegen_slab(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long B,
                    int max_rows_per_board, int8_t* afterstates52, long long afterstate_capacity_rows,
                    int8_t* row_players /*nullable*/, uint16_t* row_features_bf16 /*nullable*/,
                    int32_t* counts_true /*nullable*/, int32_t* counts, long long* starts,
                    unsigned long long* alloc_rows, int32_t* status, void* workspace, size_t workspace_bytes,
                    void* stream);

/* Tuning / test hook: threads per position of K1's overflow tiers (tier 1: 128, 256 or 512; tier 2: 512 or 1024; 0 = the
 * default choice).  Results never depend on it: the race tests run the heavy fixtures under every setting and require
 * bit-identical output.  Process-wide; not meant for production callers. */
int bg_set_team_threads(int mid, int big);

/* ------------------------------------------------------------------------------------------------
 * K3  feature encoding.  Replaces get_board_features_batch_from_tensors (ai/batching.py:78-147) ==
 * ImmutableBoard.get_board_features (board/immutable_board.py:171-212).  flags[b] = player whose
 * turn flag is set (features 196/197); for afterstates that is the MOVER (ai/batching.py:72-74).
 * If flags == NULL, flag_all (0/1) applies to every row.  (K1's row_players output is the flags
 * array of a ragged afterstate buffer.)  If n_rows_dev != NULL the number of rows encoded is
 * min(B, *n_rows_dev), read on the device (K1's alloc_rows counter), so no host sync is needed.
 */
int bg_encode_f32(const int8_t* boards52, const int8_t* flags, int flag_all, long long B,
                  const unsigned long long* n_rows_dev /*nullable*/, float* out, long long ld /* even, >= 198 */,
                  void* stream);
/* bf16 rows, ld >= 200 elements and a multiple of 8; columns 198..ld-1 are written as zeros.  ld = 208 (13 tensor-core
 * K-steps; `out` 16-byte aligned) takes the fast path: feature tiles built in shared memory, bulk asynchronous stores. */
int bg_encode_bf16(const int8_t* boards52, const int8_t* flags, int flag_all, long long B,
                   const unsigned long long* n_rows_dev /*nullable*/, uint16_t* out, long long ld, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  fused step / reward / terminal / auto-reset / dice.  Replaces BackgammonEnv.step/reset/
 * roll_dice/pass_turn/check_for_gammon/check_for_backgammon (environment/backgammon_env.py:78-191,
 * 245-251,365-405) and the auto-reset of VectorizedBackgammonEnv.step (environment/vec_bg_env.py:28-49)
 * for N games.  Legal-move lists are NOT produced here: call bg_movegen_slab on (boards, players,
 * dice) afterwards (that is update_legal_moves, backgammon_env.py:198-243).
 */
typedef struct bg_env_state {
    long long n_games;
    int8_t* boards52;        /* [N][52]  current positions */
    int8_t* players;         /* [N]      player to move */
    int8_t* dice;            /* [N][2]   current roll, as rolled */
    int32_t* scores;         /* [N][2]   match points of PLAYER1/PLAYER2 (backgammon_env.py:42-45) */
    uint32_t* draws;         /* [N]      dice draws consumed by the game slot's stream */
    int8_t* match_over;      /* [N]      set when a match ended; cleared (with scores) by the next reset */
    /* legal plays of the current positions (output of bg_movegen_slab) */
    const int8_t* afterstates52;
    const long long* starts; /* [N] */
    const int32_t* counts;   /* [N] min(true count, max_legal_moves) */
    /* dice source: Philox4x32-10 keyed by seed, counter (stream_base + game index, draw); or, if
     * ext_dice != NULL, ext_dice[g][draw][2] with ext_len draws per game. */
    unsigned long long seed;
    unsigned long long stream_base;
    const int8_t* ext_dice;
    long long ext_len;
    int32_t match_length;
    /* no_auto_reset != 0: BackgammonEnv semantics instead of VectorizedBackgammonEnv's (backgammon_env.py:119-121,
     * 156-181 vs vec_bg_env.py:35-36): a win leaves the terminal position in place with the WINNER still to move
     * (the observation the reference returns), sets game_over[g], and the NEXT step of that game resets it and
     * returns reward 0, done 1 (flags bit 2).  game_over is then required.  0 = auto-reset inside the step. */
    int32_t no_auto_reset;
    int8_t* game_over;       /* [N] latch, only with no_auto_reset */
} bg_env_state;

typedef struct bg_step_out {
    float* rewards;          /* [N]  -1 invalid, 0, 1 / 1.5 / 2 (backgammon_env.py:23-28) */
    uint8_t* dones;          /* [N] */
    int8_t* info_player;     /* [N] player to move at entry (info["current_player"], :117) */
    int8_t* winner;          /* [N] -1 = none */
    int8_t* game_score;      /* [N] 0, 1, 2, 3 */
    uint8_t* flags;          /* [N] bit0 passed, bit1 invalid action, bit2 step on a finished game = reset (no_auto_reset) */
} bg_step_out;

/* reset all games (mask == NULL) or those with mask[g] != 0: initial position + opening protocol
 * (backgammon_env.py:78-113). */
int bg_env_reset(const bg_env_state* st, const uint8_t* mask, int32_t* status, void* stream);
/* one step of every game with actions[g] (int32 index into its legal plays; ignored on a pass). */
int bg_env_step(const bg_env_state* st, const int32_t* actions, const bg_step_out* out, int32_t* status, void* stream);
/* bg_random_actions + bg_env_step in ONE launch: game g plays actions[g] = mulhi(Philox(act_seed, stream_base+g, t; "ACT1"), counts[g])
 * (the same draw as bg_random_actions), written to actions_out if non-null. */
int bg_env_step_random(const bg_env_state* st, unsigned long long act_seed, uint32_t t, int32_t* actions_out /*nullable*/,
                       const bg_step_out* out, int32_t* status, void* stream);
/* actions from (pinned) host memory to the device: one plain cudaMemcpyAsync of n int32 on `stream` (the host-side
 * policy's upload of VectorizedBackgammonEnv.step(actions), vec_bg_env.py:28-33), issued from this library so that the
 * caller needs no second CUDA runtime binding. */
int bg_copy_actions_async(int32_t* actions_dev, const int32_t* host_actions, long long n, void* stream);
/* A rollout's record of a step's inputs in one launch: boards52 [N][52], players [N], legal-play counts [N] of the env state copied to
 * the caller's buffers (the slices of its [T][N] rollout buffer: what BackgammonPPOAgent.store_transition keeps, ppo_agent.py:193-204). */
int bg_record_state(const bg_env_state* st, int8_t* boards52_out, int8_t* players_out, int32_t* counts_out, void* stream);
/* uniform random policy: actions[g] = mulhi(Philox(seed, stream_base+g, t; "ACT1"), counts[g]) (0 if none). */
int bg_random_actions(const int32_t* counts, long long N, unsigned long long seed, unsigned long long stream_base,
                      uint32_t t, int32_t* actions, void* stream);

/* update_legal_moves + get_observation of N games in ONE call (environment/backgammon_env.py:193-243):
 * bg_movegen_slab, then bg_encode_bf16 of every legal play (features_bf16, nullable; rows as in afterstates52,
 * turn flag = the mover) and bg_encode_f32 of the current positions (observations_f32, nullable; flag = player to
 * move).  With side_stream != NULL (a second stream of the caller) the encoders run on it beside K1's latency-
 * bound overflow tiers: fork and join are done with events inside the call, so for the caller everything is still
 * ordered on `stream`.  side_stream == NULL: serial on `stream`.  row_players and counts are required.
 * k1_begin_event / k1_end_event (nullable cudaEvent_t): recorded on `stream` around K1's launches, for timing K1
 * inside a fused step (bench.py's roofline). */
int bg_update_legal_plays(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long N,
                          int max_rows_per_board, int8_t* afterstates52, long long afterstate_capacity_rows,
                          int8_t* row_players, int32_t* counts_true /*nullable*/, int32_t* counts, long long* starts,
                          unsigned long long* alloc_rows, int32_t* status, void* workspace, size_t workspace_bytes,
                          uint16_t* features_bf16 /*nullable*/, long long features_ld, float* observations_f32 /*nullable*/,
                          long long observations_ld, void* k1_begin_event /*nullable*/, void* k1_end_event /*nullable*/,
                          void* side_stream /*nullable*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4  MLP leaf evaluator: value head of BackgammonPolicyNetwork.forward (agent/policy_network.py:58-75)
 * v = w_v . relu(W1 x + b1) + b_v, fused with the feature encoding: input is board52 + flag, the
 * 198-wide bf16 rows are built in shared memory and multiplied on the tcgen05 tensor cores
 * (bf16 x bf16 -> f32 in TMEM); the hidden layer is reduced in the epilogue straight out of TMEM.
 *   w1_bf16: 128 x 208 bf16 as written by bg_pack_w1 -- an OPAQUE operand tile (tcgen05 K-major no-swizzle layout, so
 *   that a kernel loads it with a straight copy): fc1.weight in the kernels' internal K order (points first,
 *   bar/off/flags last), then (fc1_bias != NULL) the bias as a bf16 hi/lo pair in columns 198, 199
 *   -- the kernels put 1.0 there in the A tile, so the bias comes out of the GEMM and b1 must then be passed as NULL --
 *   then zeros.  With fc1_bias == NULL those columns are zero and b1 is added explicitly.
 *   flags / flag_all as in K3; flip_flags = 1 evaluates every row with the OTHER player's flag.
 *   terminal_aware = 1: a row whose flag player has borne off 15 men gets the win reward 1 / 1.5 / 2
 *   (environment/backgammon_env.py:156-171) instead of the network value (2-ply leaf rule).
 *   n_rows_dev as in K3.
 */
int bg_pack_w1(const float* fc1_weight /*[128][198] f32*/, const float* fc1_bias /*[128], nullable*/,
               uint16_t* w1_bf16 /*[128][208]*/, void* stream);
int bg_mlp_value(const int8_t* boards52, const int8_t* flags, int flag_all, int flip_flags, long long B,
                 const unsigned long long* n_rows_dev /*nullable*/, const uint16_t* w1_bf16,
                 const float* b1 /*[128]; NULL = folded into w1_bf16*/, const float* wv /*[128]*/, float bv, int terminal_aware,
                 float* values, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N1  policy/value forward of the PPO rollout: BackgammonPPOAgent.select_action (agent/ppo_agent.py:138-191) =
 * BackgammonPolicyNetwork.forward (agent/policy_network.py:58-75) + log(mask + 1e-45) + softmax +
 * Categorical.sample / argmax + log_prob, for B positions given as board52 + turn flag, in one kernel: both
 * GEMMs (198->128, 128->500) on the tcgen05 tensor cores, logits reduced on chip (never written unless
 * logits_out != NULL).  The mask is the env's prefix mask: slot k is legal iff k < legal_counts[b]
 * (environment/backgammon_env.py:228-231); legal_counts == NULL = all 500 legal.  With no legal slot (a pass) the
 * reference samples from all 500 slots; so does this.
 *   wa_bf16: 512 x 128 bf16 as written by bg_pack_wa (opaque: action_head.weight, rows 500..511 zero, in the tcgen05
 *   operand layout), ba: [500] f32;
 *   w1_bf16 / b1 as in bg_mlp_value (b1 == NULL: bias folded into w1_bf16).
 *   sampling: exact categorical sampling (inverse CDF inside blocks of slots, blocks merged reservoir-style) with uniforms from
 *   Philox4x32-10 keyed by seed, counter (stream_base + b, step, block): reproducible
 *   and independent of the batch split; greedy = 1 takes argmax (lowest slot on ties) like the inference mode.
 *   outputs: actions [B] i32, log_probs [B] f32 (nullable), values [B] f32 (nullable),
 *            logits_out [B][500] f32 (nullable; unmasked logits, for parity checks).
 */
int bg_pack_wa(const float* action_head_weight /*[500][128] f32*/, uint16_t* wa_bf16 /*[512][128]*/, void* stream);
size_t bg_policy_workspace_bytes(long long B);
/* workspace (nullable): bg_policy_workspace_bytes(B) bytes of device scratch; with it (and a mask, and logits_out == NULL)
 * the rows are split into those with 1..128 legal slots, which need one quarter of the policy GEMM and only the legal
 * part of its epilogue, and the rest (passes, > 128 slots).  Results do not depend on it beyond rounding (illegal slots
 * are then left out of the softmax instead of entering it with exp(-103)). */
int bg_policy_sample(const int8_t* boards52, const int8_t* flags, int flag_all, long long B,
                     const int32_t* legal_counts /*nullable*/, const uint16_t* w1_bf16, const float* b1,
                     const uint16_t* wa_bf16, const float* ba, const float* wv, float bv, unsigned long long seed,
                     unsigned long long stream_base, uint32_t step, int greedy, int32_t* actions,
                     float* log_probs /*nullable*/, float* values /*nullable*/, float* logits_out /*nullable*/,
                     void* workspace /*nullable*/, size_t workspace_bytes, void* stream);

/* N2  discounted returns / GAE(lambda) per game over a rollout stored step-major [T][N]; replaces
 * BackgammonPPOAgent.compute_returns (agent/ppo_agent.py:206-216).  values / last_values nullable (= 0):
 * lambda = 1, last_values = NULL is compute_returns for each game; T = T*N, N = 1 is the reference's walk over its
 * interleaved memory (agent/train.py:64-66).  returns / advantages: either may be NULL. */
int bg_gae(const float* rewards, const uint8_t* dones, const float* values /*nullable*/, const float* last_values /*nullable*/,
           int T, long long N, float gamma, float lambda, float* returns, float* advantages, void* stream);

/* N2  loss of one PPO epoch (agent/ppo_agent.py:271-299: log(mask + 1e-45), softmax, log_prob, entropy, ratio, clipped
 * surrogate, MSE, their weighted sum) and its gradient w.r.t. the network outputs, one pass over the logits:
 *   logits / dlogits: [B][ld] bf16 (flags & BG_LOSS_LOGITS_BF16) or f32, ld >= 500 and a multiple of 4; values, returns, advantages,
 *   old_log_probs: [B] f32; counts: legal slots per sample (prefix mask); actions: taken slot.
 *   dlogits = d loss / d logits (columns 500 .. ld-1 are written as zeros), dvalues = d loss / d values (both already
 *   divided by B); dbias[0..min(ld,512)-1] += column sums of dlogits = d loss / d action_head.bias (nullable; caller zeroes).
 *   values == NULL: the value head was computed as row 500 of the action head's GEMM -- the value is read from column 500
 *   of the logits (ld >= 504), d loss / d value is written to column 500 of dlogits (and to dvalues if non-null) and its
 *   sum to dbias[500], so the value head's backward rides through the action head's backward GEMMs.
 *   sums[0..2] += sum_i policy term, sum_i (v - R)^2, sum_i entropy  (caller zeroes; loss = (s0 + c_v s1 - c_e s2) / B).
 *   Two launches on `stream`: rows with 1..128 legal slots and the action inside the prefix, four per warp; then the rest. */
#define BG_LOSS_LOGITS_BF16 1          /* logits / dlogits are bf16 (else f32) */
#define BG_LOSS_DLOGITS_PREZEROED 2    /* dlogits already holds zeros wherever this call would write zeros for a row with 1..128 legal slots
                                         (columns 128 .. ld-1 except the value column): they are not written again.  For a buffer that is
                                         reused with the SAME rows (the epochs of one update), zeroed once. */
int bg_ppo_loss_grad(const void* logits, int flags /* BG_LOSS_* */, long long ld, const float* values, const int32_t* counts,
                     const int32_t* actions, const float* old_log_probs, const float* advantages, const float* returns,
                     long long B, float eps_clip, float value_coef, float entropy_coef, void* dlogits, float* dvalues,
                     float* dbias /*nullable*/, float* sums, void* stream);

/* N2  the linear algebra of one PPO epoch (agent/ppo_agent.py:268-305: forward of both layers, loss.backward()) as
 * hand-written tcgen05 GEMMs (csrc/ppo_gemm.cu) -- no library GEMM on the update path.
 * Layout: every activation matrix is bf16 and TILE-BLOCKED -- tiles of 128 rows, inside a tile 16-byte chunks ordered
 * [column / 8][row][8 columns], i.e. element (r, c) of a matrix with C columns sits at 16-bit index
 * (r / 128) * (128 C) + (c / 8) * 1024 + (r % 128) * 8 + c % 8: a tile is the tensor cores' shared-memory operand image, so
 * a pipeline stage is one bulk copy.  The batch is sorted by class and each class padded to whole tiles with zero rows:
 * tiles [0, TA) = class A (1..128 legal slots, stored action among them), tiles [TA, TA + TB) = class B (passes, more slots).
 * Buffers: x [(TA+TB) tiles][208 columns] (K3's rows, column 198 = 1.0: the bias), h / dpre [..][128]; logits / dlogits of class A
 * [TA tiles][144] (slots 0..127, value head in column 128), of class B [TB tiles][512] (value head in column 500).  Tile ranges
 * are absolute tiles of x / h / dpre; the class B buffers are indexed from their own tile 0, so pass them offset by -TA tiles.
 *   bg_ppo_gather_block: x_blocked row p = x_rowmajor[perm[p]] (zero row where perm[p] < 0), column set_one_col set to 1.0.
 *   bg_ppo_pack_weights: flat f32 master weights (fc1.weight, fc1.bias, action_head.weight, action_head.bias,
 *     value_head.weight, value_head.bias: 90,101 floats) -> bf16 operand tiles w1p [26][128][8], wap_a [16][144][8],
 *     wap_b [16][512][8] and the f32 bias rows bias_a [144], bias_b [512].
 *   bg_ppo_gemm_nt(op): out = epilogue(A . W^T) for tiles [tile_begin, tile_end):
 *     HIDDEN   h = relu(x w1p^T);  LOGITS_A / _B  logits = h wap^T + bias;  DPRE_A / _B  dpre = (dlogits wap) * [h > 0]
 *     (W = the matching tile; bias for the LOGITS ops; h_mask = h for the DPRE ops).
 *   bg_ppo_gemm_tn(op): flat_grad += A^T . B over tiles [tile_begin, tile_end) (f32 atomics; caller zeroes flat_grad):
 *     GRAD_WA_A  A = h, B = dlogits_a -> action_head.weight[0..127], value_head.weight;  GRAD_WA_B  A = h, B = dlogits_b
 *     -> action_head.weight, value_head.weight;  GRAD_W1  A = dpre, B = x -> fc1.weight, fc1.bias.
 *   bg_ppo_loss_grad_classes: bg_ppo_loss_grad on the two (tile-blocked) class buffers: n_a / n_b real rows, the per-sample
 *     arrays hold class B at [b_offset, ..) (b_offset = 128 TA); means over n_a + n_b; dbias [512]: columns 0..499 =
 *     action_head.bias, 500 = value_head.bias.
 *   bg_adam_step: torch.optim.Adam's update (ppo_agent.py:83) of the flat parameters, one kernel; step counts from 1;
 *     grads are multiplied by grad_scale first (1 / world size after a summing all-reduce). */
#define BG_PPO_OP_HIDDEN 0
#define BG_PPO_OP_LOGITS_A 1
#define BG_PPO_OP_LOGITS_B 2
#define BG_PPO_OP_DPRE_A 3
#define BG_PPO_OP_DPRE_B 4
#define BG_PPO_OP_GRAD_WA_A 5
#define BG_PPO_OP_GRAD_WA_B 6
#define BG_PPO_OP_GRAD_W1 7
#define BG_PPO_NUM_PARAMS 90101
int bg_ppo_pack_weights(const float* flat_params, uint16_t* w1p, uint16_t* wap_a, uint16_t* wap_b, float* bias_a, float* bias_b,
                        void* stream);
int bg_ppo_gather_block(const uint16_t* x_rowmajor, long long ld_src, const int32_t* perm, long long rows_pad, int ncols,
                        int set_one_col /* -1: none */, uint16_t* x_blocked, void* stream);
/* The same result without the row-major features: x_blocked = tile-blocked bf16 features (208 columns, reference feature order,
 * zero padding, column set_one_col = 1.0 if >= 0) of boards52[perm[p]] with the turn flag of flags[perm[p]], zero rows where
 * perm[p] < 0 -- K3 and the gather in one pass (what PPOTrainer's update uses: the rollout stores boards, not features). */
int bg_ppo_encode_block(const int8_t* boards52, const int8_t* flags, const int32_t* perm, long long rows_pad, int set_one_col,
                        uint16_t* x_blocked, void* stream);
int bg_ppo_gemm_nt(int op, const uint16_t* A, long long tile_begin, long long tile_end, const uint16_t* W, const float* bias,
                   const uint16_t* h_mask, uint16_t* out, void* stream);
int bg_ppo_gemm_tn(int op, const uint16_t* A, const uint16_t* B, long long tile_begin, long long tile_end, float* flat_grad,
                   float* scratch /* GRAD_W1: [199][128] f32 (dW1p transposed; fc1.* of flat_grad are then overwritten) */, void* stream);
int bg_ppo_loss_grad_classes(const void* logits_a, void* dlogits_a, const void* logits_b, void* dlogits_b, long long n_a, long long n_b,
                             long long b_offset, const int32_t* counts, const int32_t* actions, const float* old_log_probs, const float* advantages,
                             const float* returns, float eps_clip, float value_coef, float entropy_coef, float* dbias /*[512]*/,
                             float* sums /*[3]*/, int class_a_done /* class A already handled by bg_ppo_logits_loss_a */, void* stream);
/* LOGITS_A with the class A loss as its epilogue (the logits of class A never leave the SM): dlogits_a = d loss / d logits of rows
 * [0, n_a) in the 144-column blocked layout (padding rows of the last tile written as zeros); dbias / sums accumulated as by
 * bg_ppo_loss_grad_classes (call that with class_a_done = 1 for class B); means over B_norm samples. */
int bg_ppo_logits_loss_a(const uint16_t* h, long long n_a, long long B_norm, const uint16_t* wap_a, const float* bias_a,
                         const int32_t* counts, const int32_t* actions, const float* old_log_probs, const float* advantages,
                         const float* returns, float eps_clip, float value_coef, float entropy_coef, uint16_t* dlogits_a, float* dbias,
                         float* sums, void* stream);
int bg_ppo_gemm_debug(int flags);  /* experiment switches for scripts/exp_ppo_gemm.py (1 no MMAs, 2 no stores, 4 no loads); 0 = normal */
int bg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                 float beta2, float eps, int step, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5  2-ply search (the reference's own 2-ply, moves/expect_minmax.py:1-206, is
 * commented-out code, so the definition is the build's, on the reference's live primitives).
 * bg_twoply (below) is the search; the pieces of the unfused pipeline stay exported (they are what the fused kernel is
 * tested against): K1 on the roots -> bg_movegen_replies_slab on the root afterstates -> bg_mlp_value on the replies
 * (terminal_aware) and on the afterstates (pass value) -> bg_twoply_scores -> bg_segment_argmax.
 */
/* replies of the OPPONENT of movers[i] to each of M positions for each of the 21 sorted rolls
 * (moves/get_all_dice_rolls.py:5-34 order): work item i*21+r; counts/starts/counts_true have M*21
 * entries; workspace must hold bg_movegen_workspace_bytes(M*21). Otherwise as bg_movegen_slab. */
int bg_movegen_replies_slab(const int8_t* positions52, const int8_t* movers, long long M,
                            int max_rows_per_board, int8_t* replies52, long long reply_capacity_rows,
                            int8_t* row_players /*nullable*/, uint16_t* row_features_bf16 /*nullable*/,
                            int32_t* counts_true /*nullable*/, int32_t* counts, long long* starts, unsigned long long* alloc_rows, int32_t* status, void* workspace,
                            size_t workspace_bytes, void* stream);
/* The inner step of the search for M root afterstates in ONE call: bg_movegen_replies_slab, bg_mlp_value on the
 * replies (terminal_aware, flag = the replying player) -> leaf_values[row], and bg_mlp_value on the positions with the
 * flag flipped -> pass_values[i].  With side_stream != NULL the leaf evaluation of the rows that are final after K1's
 * tier 0 runs beside the latency-bound overflow tiers (fork/join inside, as in bg_update_legal_plays). */
int bg_twoply_replies_values(const int8_t* positions52, const int8_t* movers, long long M, int8_t* replies52,
                             long long reply_capacity_rows, int8_t* row_players, int32_t* counts, long long* starts,
                             unsigned long long* alloc_rows, int32_t* status, void* workspace, size_t workspace_bytes,
                             const uint16_t* w1_bf16, const float* b1, const float* wv, float bv, float* leaf_values,
                             float* pass_values, void* side_stream /*nullable*/, void* stream);
/* scores[i] = +win reward if movers[i] has borne off 15 in after52[i], else
 * -sum_r p_r * (max over the replies of (i,r) of leaf_values, or pass_values[i] when there is no reply) */
int bg_twoply_scores(const float* leaf_values, const long long* reply_starts, const int32_t* reply_counts,
                     const float* pass_values, const int8_t* after52, const int8_t* movers, long long M,
                     float* scores, void* stream);
/* The whole search in ONE call, no host synchronisation (intent of the reference:
 * moves/expect_minmax.py:184 expectiminimax(board, depth 2, ...), dead code there): K1 on the B roots (slab form) ->
 * twoply_fused_kernel (csrc/twoply_fused.cu): every root afterstate decoded once, its replies to the 21 rolls generated
 * in shared memory, expanded into feature rows in TENSOR MEMORY, multiplied on the tcgen05 tensor cores and max-reduced
 * per (afterstate, roll) on chip -- no reply row, feature row or leaf value crosses HBM -- -> scores -> best play.
 * The few (afterstate, roll) items too large for the on-chip scratch go through K1's team tiers + K4 inside the call.
 *   outputs (device, caller-owned): afterstates52 [max_afterstates][52] (rows of root b at starts[b], reference
 *   legal_moves order), scores [max_afterstates] (same rows), counts [B], starts [B], best [B] (index into the root's
 *   plays, lowest on ties, -1 = no legal play), best_score [B] (nullable), stats [4] (nullable: root afterstates, leaves
 *   evaluated, (afterstate, roll) items sent through the overflow path, leaves of those).  status: BG_STATUS_OUTPUT_OVERFLOW if the roots have more than max_afterstates plays in total or the
 *   overflow slab (64 rows per afterstate + 262,144) was too small -- results are then incomplete, call again with more room.
 *   workspace: bg_twoply_workspace_bytes(B, max_afterstates) bytes; bg_workspace_bytes(BG_WS_TWOPLY, B) sizes it for
 *   max_afterstates = BG_TWOPLY_DEFAULT_ROWS_PER_ROOT * B + 4096 (the mean is 18.6 plays per position). */
#define BG_WS_MOVEGEN 0                 /* = bg_movegen_workspace_bytes(B) */
#define BG_WS_POLICY 1                  /* = bg_policy_workspace_bytes(B) */
#define BG_WS_TWOPLY 2
#define BG_TWOPLY_DEFAULT_ROWS_PER_ROOT 64
size_t bg_workspace_bytes(int kind, long long B);
size_t bg_twoply_workspace_bytes(long long B, long long max_afterstates);
int bg_twoply(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long B,
              const uint16_t* w1_bf16 /* bg_pack_w1 with the bias folded in */, const float* wv, float bv,
              long long max_afterstates, int8_t* afterstates52, float* scores, int32_t* counts, long long* starts,
              int32_t* best, float* best_score /*nullable*/, unsigned long long* stats /*nullable*/, int32_t* status,
              void* workspace, size_t workspace_bytes, void* stream);

/* best[b] = lowest index (within block b) of the maximum of scores[starts[b] .. +counts[b]), -1 if empty */
int bg_segment_argmax(const float* scores, const long long* starts, const int32_t* counts, long long B,
                      int32_t* best, float* best_score /*nullable*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BG_B200_H */
