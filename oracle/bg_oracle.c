/*
 * bg_oracle.c -- CPU restatement of the reference's backgammon hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / reported CPU baseline.
 * The product path (mlp-ppo-2ply-p3_b200/) never links, imports or calls it.
 *
 * Parity pin: the reference (Nick-qsv/MLP-PPO-2PLY-P3) ships no tests or golden
 * vectors of its own.  This restatement is pinned instead
 * against outputs of the reference itself, generated in the build container by
 * tests/golden/make_golden.py (imports /root/reference) and committed under
 * tests/golden/ -- see tests/test_oracle_golden.py.  The 2-ply search is dead
 * code in the reference (src/moves/expect_minmax.py:1-206 is commented out), so
 * bg_twoply() follows the definition in SURVEY.md section 8(c) built on the live
 * primitives; for that function parity is "unpinned by the reference" and is
 * pinned only against a Python restatement on the reference's own primitives.
 *
 * Board layout = the reference's: int8[4][24]
 *   row 0/1: checkers of PLAYER1/PLAYER2 on points 0..23
 *   row 2, col 0/1: bar of PLAYER1/PLAYER2;  row 3, col 0/1: borne off
 *   (src/board/immutable_board.py:20-27)
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define NPTS 24
#define POS_BAR 24      /* moves/move_types.py:34 */
#define POS_OFF 25      /* moves/move_types.py:35 */

enum { ST_NORMAL = 0, ST_ON_BAR = 1, ST_BEAR_OFF = 2, ST_GAME_OVER = 3 }; /* board/board_state.py:6-10 */

typedef struct { int8_t t[4][NPTS]; } Board;                  /* 96 bytes */
typedef struct { int8_t start, end, hit; } SubMove;           /* moves/move_types.py:38-42 */
typedef struct { int32_t n_sub; SubMove sub[4]; Board after; } FullMove;

/* ---------------------------------------------------------------- conditions.py */

/* conditions.py:96-108 */
static int check_for_win(const Board *b, int p) { return b->t[3][p] == 15; }
/* conditions.py:81-93 */
static int check_for_bar(const Board *b, int p) { return b->t[2][p] > 0; }
/* conditions.py:111-147 */
static int all_checkers_home(const Board *b, int p) {
    int lo = (p == 1) ? 0 : 18, hi = (p == 1) ? 6 : 24;       /* :123-126 */
    int total = 0;
    for (int idx = 0; idx < NPTS; ++idx) {                    /* :131-137 */
        int n = b->t[p][idx];
        if (n > 0) {
            if (idx >= lo && idx < hi) total += n; else return 0;
        }
    }
    if (b->t[2][p] > 0) return 0;                             /* :140-141 */
    return total + b->t[3][p] == 15;                          /* :144-147 */
}
/* conditions.py:7-33 (destination on the board only; callers never pass BEAR_OFF) */
static int valid_move(const Board *b, int p, int dest) {
    if (dest >= 0 && dest < NPTS) return b->t[1 - p][dest] < 2;
    return dest == POS_OFF;
}
/* conditions.py:36-54 */
static int check_if_blot(const Board *b, int p, int idx) {
    return idx >= 0 && idx < NPTS && b->t[1 - p][idx] == 1;
}

/* move_logic.py:258-275 */
static int compute_board_state(const Board *b, int p) {
    if (check_for_win(b, p)) return ST_GAME_OVER;
    if (check_for_bar(b, p)) return ST_ON_BAR;
    if (all_checkers_home(b, p)) return ST_BEAR_OFF;
    return ST_NORMAL;
}

/* ---------------------------------------------------------------- move_logic.py */

/* move_logic.py:47-92 */
static int get_moves_normal(const Board *b, int die, int p, SubMove *out) {
    int n = 0, dir = (p == 0) ? 1 : -1;
    for (int idx = 0; idx < NPTS; ++idx) {                    /* ascending for BOTH players, :67 */
        if (b->t[p][idx] > 0) {
            int dest = idx + die * dir;
            if (dest >= 0 && dest < NPTS && valid_move(b, p, dest)) {
                out[n].start = (int8_t)idx; out[n].end = (int8_t)dest;
                out[n].hit = (int8_t)check_if_blot(b, p, dest); ++n;
            }
        }
    }
    return n;
}
/* move_logic.py:95-137 */
static int get_moves_bar(const Board *b, int die, int p, SubMove *out) {
    int dest = (p == 0) ? die - 1 : 24 - die;                 /* :111-114 */
    int lo = (p == 0) ? 0 : 18, hi = (p == 0) ? 6 : 24;       /* :117-120 */
    if (dest >= lo && dest < hi) {
        if (dest >= 0 && dest < 24 && b->t[1 - p][dest] < 2) {/* is_valid_entry_at_index, conditions.py:57-78 */
            out[0].start = POS_BAR; out[0].end = (int8_t)dest;
            out[0].hit = (int8_t)check_if_blot(b, p, dest);
            return 1;
        }
    }
    return 0;
}
/* move_logic.py:140-255 */
static int get_moves_bear_off(const Board *b, int die, int p, SubMove *out) {
    int n = 0;
    int h0 = (p == 0) ? 18 : 0, dir = (p == 0) ? 1 : -1;
    int last = (p == 0) ? 18 : 5;                             /* :165,:169 */
    for (int idx = h0; idx < h0 + 6; ++idx) {                 /* 1. in-home normal moves, :172-193 */
        if (b->t[p][idx] > 0) {
            int dest = idx + die * dir;
            if (dest >= 0 && dest < NPTS && valid_move(b, p, dest)) {
                out[n].start = (int8_t)idx; out[n].end = (int8_t)dest;
                out[n].hit = (int8_t)check_if_blot(b, p, dest); ++n;
            }
        }
    }
    if (p == 0) {                                             /* 2. farthest checker, :196-207 */
        for (int idx = 18; idx < 24; ++idx) if (b->t[p][idx] > 0) { last = idx; break; }
    } else {
        for (int idx = 5; idx >= 0; --idx) if (b->t[p][idx] > 0) { last = idx; break; }
    }
    if (p == 0) {                                             /* 3. bear-off moves, :210-231 */
        if (last + die >= NPTS) { out[n].start = (int8_t)last; out[n].end = POS_OFF; out[n].hit = 0; ++n; }
        int ps = NPTS - die;
        if (ps != last && ps >= 18 && ps < 24 && b->t[p][ps] > 0) {
            out[n].start = (int8_t)ps; out[n].end = POS_OFF; out[n].hit = 0; ++n;
        }
    } else {                                                  /* :232-253 */
        if (last - die < 0) { out[n].start = (int8_t)last; out[n].end = POS_OFF; out[n].hit = 0; ++n; }
        int ps = die - 1;
        if (ps != last && ps >= 0 && ps < 6 && b->t[p][ps] > 0) {
            out[n].start = (int8_t)ps; out[n].end = POS_OFF; out[n].hit = 0; ++n;
        }
    }
    return n;
}
/* move_logic.py:20-44 ; at most 15 occupied points + 1 extra bear-off => <= 16 */
#define MAX_SUB 32
static int get_moves_with_one_die(const Board *b, int die, int p, SubMove *out) {
    switch (compute_board_state(b, p)) {
        case ST_NORMAL:   return get_moves_normal(b, die, p, out);
        case ST_ON_BAR:   return get_moves_bar(b, die, p, out);
        case ST_BEAR_OFF: return get_moves_bear_off(b, die, p, out);
        default:          return 0;
    }
}

/* board/immutable_board.py:42-89 (guard branches return the board unchanged) */
static Board move_checker(const Board *b, int p, SubMove m) {
    Board nb = *b;
    int o = 1 - p;
    if (m.start == POS_BAR) {
        if (nb.t[2][p] > 0) nb.t[2][p] -= 1; else return *b;              /* :56-62 */
    } else {
        if (nb.t[p][m.start] > 0) nb.t[p][m.start] -= 1; else return *b;  /* :64-70 */
    }
    if (m.hit) {                                                          /* :73-81 */
        if (nb.t[o][m.end] > 0) { nb.t[o][m.end] -= 1; nb.t[2][o] += 1; } else return *b;
    }
    if (m.end == POS_OFF) nb.t[3][p] += 1; else nb.t[p][m.end] += 1;      /* :84-87 */
    return nb;
}

/* ---------------------------------------------------------------- handle_moves.py */

typedef struct {
    FullMove *mv; int n, cap;
    /* unique_boards: the reference keeps a Python set of hash(96 bytes)
       (immutable_board.py:236-246, handle_moves.py:334-339); a 64-bit hash of
       distinct 96-byte strings is collision-free for all practical purposes, so
       the restatement uses exact byte equality. */
    int32_t *tab; int tabsz;
} MoveList;

static uint32_t board_hash32(const Board *b) {
    const uint8_t *q = (const uint8_t *)b; uint32_t h = 2166136261u;
    for (int i = 0; i < 96; ++i) { h ^= q[i]; h *= 16777619u; }
    return h;
}
static void ml_init(MoveList *l) {
    l->n = 0; l->cap = 64; l->mv = (FullMove *)malloc(sizeof(FullMove) * l->cap);
    l->tabsz = 256; l->tab = (int32_t *)malloc(sizeof(int32_t) * l->tabsz);
    for (int i = 0; i < l->tabsz; ++i) l->tab[i] = -1;
}
static void ml_free(MoveList *l) { free(l->mv); free(l->tab); }
static void ml_rehash(MoveList *l) {
    free(l->tab); l->tabsz *= 4; l->tab = (int32_t *)malloc(sizeof(int32_t) * l->tabsz);
    for (int i = 0; i < l->tabsz; ++i) l->tab[i] = -1;
    for (int i = 0; i < l->n; ++i) {
        uint32_t h = board_hash32(&l->mv[i].after) & (l->tabsz - 1);
        while (l->tab[h] >= 0) h = (h + 1) & (l->tabsz - 1);
        l->tab[h] = i;
    }
}
/* handle_moves.py:313-341  -- first sequence to reach a board wins */
static void add_unique_board(MoveList *l, const Board *after, const SubMove *subs, int n_sub) {
    uint32_t h = board_hash32(after) & (l->tabsz - 1);
    while (l->tab[h] >= 0) {
        if (memcmp(&l->mv[l->tab[h]].after, after, sizeof(Board)) == 0) return;
        h = (h + 1) & (l->tabsz - 1);
    }
    if (l->n == l->cap) { l->cap *= 2; l->mv = (FullMove *)realloc(l->mv, sizeof(FullMove) * l->cap); }
    FullMove *f = &l->mv[l->n];
    f->n_sub = n_sub; memset(f->sub, 0, sizeof f->sub);
    for (int i = 0; i < n_sub; ++i) f->sub[i] = subs[i];
    f->after = *after;
    l->tab[h] = l->n++;
    if (l->n * 2 > l->tabsz) ml_rehash(l);
}

/* handle_moves.py:109-200 */
static void handle_non_doubles(const Board *b, int hi, int lo, MoveList *l, int p, int reverse) {
    int d0 = reverse ? lo : hi, d1 = reverse ? hi : lo;       /* :134 */
    SubMove first[MAX_SUB], second[MAX_SUB], seq[2];
    int nf = get_moves_with_one_die(b, d0, p, first);         /* :137-139 */
    int two = 0;
    for (int i = 0; i < nf && !two; ++i) {                    /* :145-155 */
        Board r = move_checker(b, p, first[i]);
        if (get_moves_with_one_die(&r, d1, p, second) > 0) two = 1;
    }
    for (int i = 0; i < nf; ++i) {                            /* :158-200 */
        Board r = move_checker(b, p, first[i]);
        int ns = get_moves_with_one_die(&r, d1, p, second);
        if (two) {
            for (int j = 0; j < ns; ++j) {
                Board r2 = move_checker(&r, p, second[j]);
                seq[0] = first[i]; seq[1] = second[j];
                add_unique_board(l, &r2, seq, 2);
            }
        } else {
            seq[0] = first[i];
            add_unique_board(l, &r, seq, 1);
        }
    }
}

/* handle_moves.py:203-310 */
static void handle_doubles(const Board *b, int die, MoveList *l, int p) {
    SubMove m1[MAX_SUB], m2[MAX_SUB], m3[MAX_SUB], m4[MAX_SUB], seq[4];
    int n1 = get_moves_with_one_die(b, die, p, m1);
    int four = 0;                                             /* :227 */
    for (int i = 0; i < n1; ++i) {
        Board b1 = move_checker(b, p, m1[i]);
        int n2 = get_moves_with_one_die(&b1, die, p, m2);
        seq[0] = m1[i];
        if (n2 == 0 && n1 && !four) add_unique_board(l, &b1, seq, 1);            /* :236-248 */
        for (int j = 0; j < n2; ++j) {
            Board b2 = move_checker(&b1, p, m2[j]);
            int n3 = get_moves_with_one_die(&b2, die, p, m3);
            seq[1] = m2[j];
            if (n3 == 0 && n2 && !four) add_unique_board(l, &b2, seq, 2);        /* :257-269 */
            for (int k = 0; k < n3; ++k) {
                Board b3 = move_checker(&b2, p, m3[k]);
                int n4 = get_moves_with_one_die(&b3, die, p, m4);
                seq[2] = m3[k];
                if (n4 == 0 && n3 && !four) add_unique_board(l, &b3, seq, 3);    /* :282-294 */
                for (int q = 0; q < n4; ++q) {
                    Board b4 = move_checker(&b3, p, m4[q]);
                    seq[3] = m4[q];
                    add_unique_board(l, &b4, seq, 4);                            /* :303-309 */
                    four = 1;                                                    /* :310 */
                }
            }
        }
    }
}

/* moves/get_all_moves.py:9-94.  Returns the plays in the reference's list order. */
static void get_all_possible_moves(int p, const Board *b, int d0, int d1, MoveList *l) {
    if (d0 != d1) {
        int hi = d0 > d1 ? d0 : d1, lo = d0 > d1 ? d1 : d0;   /* :30 */
        handle_non_doubles(b, hi, lo, l, p, 0);               /* :33-39 */
        if (l->n == 0 || !(l->n == 1 && l->mv[0].n_sub == 1)) /* :43-45  (skip-reverse shortcut, Q1) */
            handle_non_doubles(b, hi, lo, l, p, 1);           /* :46-53 */
    } else {
        handle_doubles(b, d0, l, p);                          /* :59-65 */
    }
    /* filter_full_moves_by_max_submoves, :73-94 (dedupe happened BEFORE this) */
    int mx = 0;
    for (int i = 0; i < l->n; ++i) if (l->mv[i].n_sub > mx) mx = l->mv[i].n_sub;
    int w = 0;
    for (int i = 0; i < l->n; ++i) if (l->mv[i].n_sub == mx) l->mv[w++] = l->mv[i];
    l->n = w;
}

/* ---------------------------------------------------------------- public: move generation */

/* Number of legal plays for (board, player, dice); if `after` != NULL writes up
 * to `cap` afterstates (96 bytes each, reference list order) and, if non-NULL,
 * their sub-move counts and sub-moves (start,end,hit)x4. Returns the TRUE count. */
int bg_legal_moves(const int8_t *board96, int player, int d0, int d1,
                   int8_t *after, int32_t *n_sub, int8_t *subs12, int cap) {
    MoveList l; ml_init(&l);
    get_all_possible_moves(player, (const Board *)board96, d0, d1, &l);
    int n = l.n;
    for (int i = 0; i < n && i < cap; ++i) {
        if (after) memcpy(after + 96 * (size_t)i, &l.mv[i].after, 96);
        if (n_sub) n_sub[i] = l.mv[i].n_sub;
        if (subs12) for (int k = 0; k < 4; ++k) {
            subs12[12 * i + 3 * k + 0] = l.mv[i].sub[k].start;
            subs12[12 * i + 3 * k + 1] = l.mv[i].sub[k].end;
            subs12[12 * i + 3 * k + 2] = l.mv[i].sub[k].hit;
        }
    }
    ml_free(&l);
    return n;
}

/* Batched: counts[B]; if after != NULL, rows are written at offsets[b] (rows,
 * exclusive scan of counts supplied by the caller). */
void bg_legal_moves_batch(const int8_t *boards96, const int8_t *players, const int8_t *dice2,
                          int B, int32_t *counts, const int64_t *offsets, int8_t *after) {
    for (int b = 0; b < B; ++b) {
        MoveList l; ml_init(&l);
        get_all_possible_moves(players[b], (const Board *)(boards96 + 96 * (size_t)b),
                               dice2[2 * b], dice2[2 * b + 1], &l);
        counts[b] = l.n;
        if (after && offsets)
            for (int i = 0; i < l.n; ++i) memcpy(after + 96 * (size_t)(offsets[b] + i), &l.mv[i].after, 96);
        ml_free(&l);
    }
}

/* ---------------------------------------------------------------- public: feature encoding */

/* board/immutable_board.py:171-212  ==  ai/batching.py:78-147 (bit-identical) */
void bg_encode(const int8_t *board96, int flag_player, float *f198) {
    const Board *b = (const Board *)board96;
    int fi = 0;
    for (int p = 0; p < 2; ++p) {
        for (int pt = 0; pt < NPTS; ++pt) {
            int c = b->t[p][pt];
            float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            if (c == 1) s0 = 1.0f;
            else if (c == 2) { s0 = 1.0f; s1 = 1.0f; }
            else if (c >= 3) { s0 = s1 = s2 = 1.0f; s3 = ((float)c - 3.0f) / 2.0f; }
            f198[fi] = s0; f198[fi + 1] = s1; f198[fi + 2] = s2; f198[fi + 3] = s3; fi += 4;
        }
        f198[fi++] = (float)b->t[2][p] / 2.0f;
        f198[fi++] = (float)b->t[3][p] / 15.0f;
    }
    f198[fi++] = (flag_player == 0) ? 1.0f : 0.0f;
    f198[fi++] = (flag_player == 0) ? 0.0f : 1.0f;
}
void bg_encode_batch(const int8_t *boards96, const int8_t *flags, int B, float *out) {
    for (int b = 0; b < B; ++b) bg_encode(boards96 + 96 * (size_t)b, flags[b], out + 198 * (size_t)b);
}

/* ---------------------------------------------------------------- public: terminal / reward */

/* environment/backgammon_env.py:156-171,365-405.  Board is the position right
 * after `player` moved.  Returns 0 if player has not won, else game score 1/2/3;
 * *reward gets 1.0 / 1.5 / 2.0 (backgammon_env.py:26-28). */
int bg_win_score(const int8_t *board96, int player, float *reward) {
    const Board *b = (const Board *)board96;
    if (b->t[3][player] != 15) { if (reward) *reward = 0.0f; return 0; }
    int o = 1 - player, score;
    int backgammon = 0;
    if (b->t[3][o] == 0) {                                    /* :385-387 */
        int h0 = (player == 0) ? 18 : 0;                      /* :390-393 */
        for (int idx = h0; idx < h0 + 6; ++idx) if (b->t[o][idx] > 0) backgammon = 1;  /* :396-398 */
        if (b->t[2][o] > 0) backgammon = 1;                   /* :401-403 */
    }
    if (backgammon) score = 3;
    else if (b->t[3][o] == 0) score = 2;                      /* check_for_gammon :365-373 */
    else score = 1;
    if (reward) *reward = score == 3 ? 2.0f : (score == 2 ? 1.5f : 1.0f);
    return score;
}

/* ---------------------------------------------------------------- Philox4x32-10 (published algorithm,
 * Salmon et al. SC'11) -- the engine's counter-based dice; restated here so the
 * oracle env can replay the engine's dice stream. */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* Engine dice convention (DESIGN.md "Dice"): draw index t of game stream g under
 * seed s = Philox(counter=(g_lo, g_hi, t, 0x44494345 "DICE"), key=(s_lo, s_hi));
 * die0 = 1 + mulhi(out0, 6), die1 = 1 + mulhi(out1, 6). */
void bg_philox_dice(uint64_t seed, uint64_t stream, uint32_t draw, int8_t *d2) {
    uint32_t o[4];
    philox4x32_10((uint32_t)stream, (uint32_t)(stream >> 32), draw, 0x44494345u,
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    d2[0] = (int8_t)(1 + (int)(((uint64_t)o[0] * 6u) >> 32));
    d2[1] = (int8_t)(1 + (int)(((uint64_t)o[1] * 6u) >> 32));
}
/* Engine random-policy convention: action = mulhi(out2 of the same Philox block
 * with domain tag 0x41435431 "ACT1", n). */
uint32_t bg_philox_action(uint64_t seed, uint64_t stream, uint32_t draw, uint32_t n) {
    uint32_t o[4];
    philox4x32_10((uint32_t)stream, (uint32_t)(stream >> 32), draw, 0x41435431u,
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return (uint32_t)(((uint64_t)o[0] * n) >> 32);
}

/* ---------------------------------------------------------------- public: environment */

/* State of one BackgammonEnv (environment/backgammon_env.py:38-76). */
typedef struct {
    Board board;
    int32_t current_player, game_over, match_over, match_length, max_legal_moves;
    int32_t score[2];
    int32_t roll[2];
    int32_t n_legal;            /* after truncation to max_legal_moves (:218-223) */
    int32_t n_legal_true;       /* before truncation */
    /* dice source: external list (dice != NULL) or Philox (seed, stream, draw) */
    const int8_t *dice; int64_t dice_len, dice_pos;
    uint64_t seed, stream; uint32_t draw;
    MoveList legal;
} Env;

static void env_roll(Env *e) {                                /* backgammon_env.py:245-246 */
    if (e->dice) {
        if (e->dice_pos < e->dice_len) {
            e->roll[0] = e->dice[2 * e->dice_pos]; e->roll[1] = e->dice[2 * e->dice_pos + 1];
        } else { e->roll[0] = 1; e->roll[1] = 2; }
        e->dice_pos++;
    } else {
        int8_t d[2]; bg_philox_dice(e->seed, e->stream, e->draw++, d);
        e->roll[0] = d[0]; e->roll[1] = d[1];
    }
}
static void env_update_legal(Env *e) {                        /* :198-243 */
    ml_free(&e->legal); ml_init(&e->legal);
    get_all_possible_moves(e->current_player, &e->board, e->roll[0], e->roll[1], &e->legal);
    e->n_legal_true = e->legal.n;
    if (e->legal.n > e->max_legal_moves) e->legal.n = e->max_legal_moves;   /* :218-223 */
    e->n_legal = e->legal.n;
}
static void board_initial(Board *b) {                         /* immutable_board.py:25-40 */
    memset(b, 0, sizeof *b);
    b->t[0][0] = 2; b->t[0][11] = 5; b->t[0][16] = 3; b->t[0][18] = 5;
    b->t[1][23] = 2; b->t[1][12] = 5; b->t[1][7] = 3; b->t[1][5] = 5;
}

Env *bg_env_new(int match_length, int max_legal_moves) {
    Env *e = (Env *)calloc(1, sizeof(Env));
    e->match_length = match_length; e->max_legal_moves = max_legal_moves;
    board_initial(&e->board); e->current_player = 0;          /* :51-54 */
    ml_init(&e->legal);
    return e;
}
void bg_env_free(Env *e) { ml_free(&e->legal); free(e); }
void bg_env_set_external_dice(Env *e, const int8_t *dice, int64_t n_pairs) {
    e->dice = dice; e->dice_len = n_pairs; e->dice_pos = 0;
}
void bg_env_set_philox(Env *e, uint64_t seed, uint64_t stream) {
    e->dice = NULL; e->seed = seed; e->stream = stream; e->draw = 0;
}
int64_t bg_env_dice_consumed(const Env *e) { return e->dice ? e->dice_pos : (int64_t)e->draw; }

/* backgammon_env.py:78-113 */
void bg_env_reset(Env *e) {
    if (e->match_over) { e->score[0] = e->score[1] = 0; e->match_over = 0; }     /* :79-82 */
    board_initial(&e->board); e->game_over = 0;                                   /* :85-86 */
    e->current_player = 1 - e->current_player;                                    /* :89-91 (dead store) */
    env_roll(e); while (e->roll[0] == e->roll[1]) env_roll(e);                    /* :94-96 */
    e->current_player = (e->roll[0] < e->roll[1]) ? 1 : 0;                        /* :99-102 */
    env_roll(e); while (e->roll[0] == e->roll[1]) env_roll(e);                    /* :105-107 */
    env_update_legal(e);                                                          /* :110 */
}

/* backgammon_env.py:115-191.  flags: bit0 passed, bit1 invalid, bit2 won.
 * Returns done.  *winner = -1 if none. */
int bg_env_step(Env *e, int action, float *reward, int32_t *info_player,
                int32_t *flags, int32_t *winner, int32_t *game_score) {
    *info_player = e->current_player; *flags = 0; *winner = -1; *game_score = 0;  /* :117 */
    if (e->game_over) { bg_env_reset(e); *reward = 0.0f; return 1; }              /* :119-121 */
    if (e->n_legal == 0) {                                                        /* :124-140 */
        *reward = 0.0f; *flags = 1;
        e->current_player = 1 - e->current_player; env_roll(e); env_update_legal(e);
        return 0;
    }
    if (action < 0 || action >= e->max_legal_moves || action >= e->n_legal) {     /* :143-149 */
        *reward = -1.0f; *flags = 2; return 0;
    }
    e->board = e->legal.mv[action].after;                                         /* :152-153 */
    int done = 0;
    if (e->board.t[3][e->current_player] == 15) {                                 /* :156-181 */
        int sc = bg_win_score((const int8_t *)&e->board, e->current_player, reward);
        *winner = e->current_player; *game_score = sc; *flags = 4;
        e->score[e->current_player] += sc; e->game_over = 1; done = 1;
        if (e->score[e->current_player] >= e->match_length) e->match_over = 1;
    } else {                                                                      /* :182-188 */
        *reward = 0.0f;
        e->current_player = 1 - e->current_player; env_roll(e); env_update_legal(e);
    }
    return done;
}
/* accessors */
void bg_env_get(const Env *e, int8_t *board96, int32_t *player, int32_t *roll2,
                int32_t *n_legal, int32_t *n_legal_true, int32_t *score2, int32_t *game_over, int32_t *match_over) {
    if (board96) memcpy(board96, &e->board, 96);
    if (player) *player = e->current_player;
    if (roll2) { roll2[0] = e->roll[0]; roll2[1] = e->roll[1]; }
    if (n_legal) *n_legal = e->n_legal;
    if (n_legal_true) *n_legal_true = e->n_legal_true;
    if (score2) { score2[0] = e->score[0]; score2[1] = e->score[1]; }
    if (game_over) *game_over = e->game_over;
    if (match_over) *match_over = e->match_over;
}
void bg_env_set_position(Env *e, const int8_t *board96, int player, int d0, int d1) {
    memcpy(&e->board, board96, 96); e->current_player = player;
    e->roll[0] = d0; e->roll[1] = d1; e->game_over = 0; env_update_legal(e);
}
void bg_env_get_afterstates(const Env *e, int8_t *after) {
    for (int i = 0; i < e->n_legal; ++i) memcpy(after + 96 * (size_t)i, &e->legal.mv[i].after, 96);
}
void bg_env_observation(const Env *e, float *f198) {          /* :193-196 */
    bg_encode((const int8_t *)&e->board, e->current_player, f198);
}

/* Random-vs-random rollout of `steps` vec-env steps (vec_bg_env.py:28-49 semantics:
 * auto-reset on done) with the engine's Philox conventions; used as the CPU
 * baseline and as the trajectory oracle for the engine's Philox mode.
 * If encode != 0 also does the per-step work the reference does in
 * update_legal_moves / get_observation (afterstate features + observation).
 * The caller must have called bg_env_reset() once.  Returns the number of env
 * steps executed; accumulates a checksum to keep the work live. */
int64_t bg_env_random_rollout(Env *e, int64_t steps, uint64_t act_seed, int encode, double *checksum,
                              int64_t *games_done, int64_t *passes) {
    float feat[198]; double cs = 0; int64_t gd = 0, ps = 0;
    for (int64_t s = 0; s < steps; ++s) {
        int action = 0;
        if (e->n_legal > 0) action = (int)bg_philox_action(act_seed, e->stream, (uint32_t)s, (uint32_t)e->n_legal);
        float r; int32_t ip, fl, w, gs;
        int done = bg_env_step(e, action, &r, &ip, &fl, &w, &gs);
        cs += r; if (fl & 1) ps++;
        if (done) { gd++; bg_env_reset(e); }                  /* vec_bg_env.py:35-36 */
        if (encode) {
            for (int i = 0; i < e->n_legal; ++i) {            /* generate_all_board_features, batching.py:10-75 (Q9: mover's flag) */
                bg_encode((const int8_t *)&e->legal.mv[i].after, e->current_player, feat);
                cs += feat[97] + feat[195];
            }
            bg_env_observation(e, feat); cs += feat[96];
        }
    }
    if (checksum) *checksum = cs;
    if (games_done) *games_done = gd;
    if (passes) *passes = ps;
    return steps;
}

/* ---------------------------------------------------------------- public: MLP + 2-ply */

/* agent/policy_network.py:58-75, value head only: v = w_v . relu(W1 x + b1) + b_v
 * W1 is (128,198) row-major (fc1.weight), f32, sequential accumulation. */
float bg_mlp_value(const float *x198, const float *W1, const float *b1, const float *wv, float bv, int hidden) {
    float v = bv;
    for (int h = 0; h < hidden; ++h) {
        float a = b1[h];
        const float *w = W1 + 198 * (size_t)h;
        for (int k = 0; k < 198; ++k) a += w[k] * x198[k];
        if (a > 0) v += wv[h] * a;
    }
    return v;
}
void bg_mlp_value_batch(const float *x, int B, const float *W1, const float *b1, const float *wv, float bv,
                        int hidden, float *out) {
    for (int b = 0; b < B; ++b) out[b] = bg_mlp_value(x + 198 * (size_t)b, W1, b1, wv, bv, hidden);
}
/* Same with operands rounded to bf16 (RNE) first -- used to bound the GPU's
 * bf16-input / f32-accumulate evaluator tightly. */
static float bf16_round(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u); r &= 0xFFFF0000u;
    float o; memcpy(&o, &r, 4); return o;
}
float bg_mlp_value_bf16(const float *x198, const float *W1, const float *b1, const float *wv, float bv, int hidden) {
    double v = bv;
    for (int h = 0; h < hidden; ++h) {
        double a = b1[h];
        const float *w = W1 + 198 * (size_t)h;
        for (int k = 0; k < 198; ++k) a += (double)bf16_round(w[k]) * (double)bf16_round(x198[k]);
        if (a > 0) v += (double)wv[h] * a;
    }
    return (float)v;
}

/* 21 sorted rolls and probabilities: moves/get_all_dice_rolls.py:5-34 */
static const float P1_36 = 1.0f / 36.0f, P2_36 = 2.0f / 36.0f;

/* 2-ply per SURVEY.md 8(c).  Root (board, me, d0, d1).  Writes scores[n] for the
 * n legal plays (reference order) and returns n (true count; at most cap scores
 * written).  *best = argmax (lowest index on ties), -1 if n == 0.
 * leaf_count (optional) accumulates the number of MLP evaluations.
 * use_bf16: evaluate leaves with bf16-rounded inputs/weights. */
int bg_twoply(const int8_t *board96, int me, int d0, int d1,
              const float *W1, const float *b1, const float *wv, float bv, int hidden,
              float *scores, int cap, int32_t *best, int64_t *leaf_count, int use_bf16) {
    MoveList A; ml_init(&A);
    get_all_possible_moves(me, (const Board *)board96, d0, d1, &A);
    int opp = 1 - me; float feat[198];
    int bi = -1; float bs = 0;
    for (int i = 0; i < A.n; ++i) {
        const Board *Ai = &A.mv[i].after;
        float score;
        if (Ai->t[3][me] == 15) {
            bg_win_score((const int8_t *)Ai, me, &score);
        } else {
            float acc = 0.0f;
            for (int r0 = 1; r0 <= 6; ++r0) for (int r1 = r0; r1 <= 6; ++r1) {
                float pr = (r0 == r1) ? P1_36 : P2_36;
                MoveList R; ml_init(&R);
                get_all_possible_moves(opp, Ai, r0, r1, &R);
                float vr;
                if (R.n == 0) {
                    bg_encode((const int8_t *)Ai, opp, feat);
                    vr = use_bf16 ? bg_mlp_value_bf16(feat, W1, b1, wv, bv, hidden)
                                  : bg_mlp_value(feat, W1, b1, wv, bv, hidden);
                    if (leaf_count) (*leaf_count)++;
                } else {
                    vr = -INFINITY;
                    for (int j = 0; j < R.n; ++j) {
                        const Board *Bj = &R.mv[j].after; float lv;
                        if (Bj->t[3][opp] == 15) bg_win_score((const int8_t *)Bj, opp, &lv);
                        else {
                            bg_encode((const int8_t *)Bj, opp, feat);
                            lv = use_bf16 ? bg_mlp_value_bf16(feat, W1, b1, wv, bv, hidden)
                                          : bg_mlp_value(feat, W1, b1, wv, bv, hidden);
                            if (leaf_count) (*leaf_count)++;
                        }
                        if (lv > vr) vr = lv;
                    }
                }
                acc += pr * vr;
                ml_free(&R);
            }
            score = -acc;
        }
        if (i < cap && scores) scores[i] = score;
        if (bi < 0 || score > bs) { bi = i; bs = score; }
    }
    if (best) *best = bi;
    int n = A.n; ml_free(&A);
    return n;
}

/* ---------------------------------------------------------------- packed 52-byte layout helpers
 * Engine HBM layout: [P1 points 24][P2 points 24][bar1 bar2 off1 off2]. */
void bg_pack52(const int8_t *board96, int B, int8_t *out52) {
    for (int b = 0; b < B; ++b) {
        const int8_t *s = board96 + 96 * (size_t)b; int8_t *d = out52 + 52 * (size_t)b;
        memcpy(d, s, 48); d[48] = s[48]; d[49] = s[49]; d[50] = s[72]; d[51] = s[73];
    }
}
void bg_unpack52(const int8_t *in52, int B, int8_t *board96) {
    for (int b = 0; b < B; ++b) {
        const int8_t *s = in52 + 52 * (size_t)b; int8_t *d = board96 + 96 * (size_t)b;
        memset(d, 0, 96); memcpy(d, s, 48); d[48] = s[48]; d[49] = s[49]; d[72] = s[50]; d[73] = s[51];
    }
}
