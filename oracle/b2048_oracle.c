/*
 * b2048_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the
 * reference's game rules and environment step, used as the parity checker for
 * the CUDA path and as the native CPU baseline in bench.py.  Nothing in the
 * product package links or calls this file.
 *
 * Parity pin: the reference ships no tests / golden vectors, so this oracle is
 * pinned against the LIVE reference (imported from /root/reference in the build
 * container by tests/golden/gen_golden.py) and against the fixtures that script
 * committed under tests/golden/ (all 65,536 rows of _row_move_left, random
 * board x action moves, full replayed episodes).  See tests/test_oracle_*.py.
 *
 * Deliberately written cell-by-cell on unpacked 4x4 arrays (no LUT, no bit
 * tricks) so that it is an independent statement of the rules, not a copy of
 * the kernel's method.  Each function cites the reference lines it follows
 * (paths relative to /root/reference).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <stdint.h>
#include <string.h>
#include <math.h>

#include "../include/b2048.h"

/* ---------------------------------------------------------------- packing */

static void unpack(uint64_t b, int cell[4][4]) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            cell[r][c] = (int)((b >> (4 * (4 * r + c))) & 0xF);
}

static uint64_t pack(int cell[4][4]) {
    uint64_t b = 0;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            b |= (uint64_t)(cell[r][c] & 0xF) << (4 * (4 * r + c));
    return b;
}

/* ------------------------------------------------------------- row move */

/* Game2048._row_move_left (src/game2048.py:120-137) on exponents: drop zeros,
 * one left-to-right pass merging equal neighbours once (value<<1 == exponent+1).
 * merged[] receives the exponents of the new tiles in emission order.
 * A 15+15 merge (exponent 16) is outside the 4-bit domain: the cell saturates
 * at 15, merged records 16 and *overflow is set. */
static int line_move(const int in[4], int out[4], int merged[2], int* overflow) {
    int cells[4], n = 0, nm = 0, w = 0;
    for (int i = 0; i < 4; ++i)
        if (in[i] != 0) cells[n++] = in[i];
    for (int i = 0; i < 4; ++i) out[i] = 0;
    int i = 0;
    while (i < n) {
        if (i + 1 < n && cells[i] == cells[i + 1]) {
            int e = cells[i] + 1;
            if (e > 15) { *overflow = 1; out[w] = 15; }
            else out[w] = e;
            merged[nm++] = e;
            i += 2;
        } else {
            out[w] = cells[i];
            i += 1;
        }
        ++w;
    }
    return nm;
}

/* Row table entry for one 16-bit row (nibble c = cell c): used to pin the
 * device LUT.  merge byte: low nibble = first merged exponent, high nibble =
 * second; 0 = none, 1 = the out-of-domain exponent 16. */
void orc_row_move_left(uint16_t row, uint16_t* out_row, uint8_t* merge_byte, int32_t* score) {
    int in[4], out[4], merged[2] = {0, 0}, ov = 0;
    for (int c = 0; c < 4; ++c) in[c] = (row >> (4 * c)) & 0xF;
    int nm = line_move(in, out, merged, &ov);
    uint16_t o = 0;
    for (int c = 0; c < 4; ++c) o |= (uint16_t)(out[c] << (4 * c));
    uint8_t mb = 0;
    int32_t sc = 0;
    for (int k = 0; k < nm; ++k) {
        int code = merged[k] == 16 ? 1 : merged[k];
        mb |= (uint8_t)(code << (4 * k));
        sc += (int32_t)1 << merged[k];
    }
    if (out_row) *out_row = o;
    if (merge_byte) *merge_byte = mb;
    if (score) *score = sc;
}

/* --------------------------------------------------------------- move */

/* Game2048._move (src/game2048.py:158-165): rotate so the move direction
 * becomes "left", slide every row, rotate back.  Restated per line: index 0 of
 * each line is the wall the tiles move toward.  Action map (game2048.py:9):
 * 0 up, 1 right, 2 down, 3 left.  merged_out collects up to 8 exponents. */
static int board_move(int cell[4][4], int action, int merged_out[8], int* n_merged, int* overflow) {
    int changed = 0;
    *n_merged = 0;
    for (int l = 0; l < 4; ++l) {
        int in[4], out[4], merged[2];
        for (int i = 0; i < 4; ++i) {
            switch (action) {
                case 3: in[i] = cell[l][i]; break;        /* left : row l, from col 0   */
                case 1: in[i] = cell[l][3 - i]; break;    /* right: row l, from col 3   */
                case 0: in[i] = cell[i][l]; break;        /* up   : col l, from row 0   */
                default: in[i] = cell[3 - i][l]; break;   /* down : col l, from row 3   */
            }
        }
        int nm = line_move(in, out, merged, overflow);
        for (int k = 0; k < nm; ++k) merged_out[(*n_merged)++] = merged[k];
        for (int i = 0; i < 4; ++i) {
            if (out[i] != in[i]) changed = 1;
            switch (action) {
                case 3: cell[l][i] = out[i]; break;
                case 1: cell[l][3 - i] = out[i]; break;
                case 0: cell[i][l] = out[i]; break;
                default: cell[3 - i][l] = out[i]; break;
            }
        }
    }
    return changed;
}

/* Game2048._is_done (src/game2048.py:172-187) */
static int is_done(int cell[4][4]) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            if (cell[r][c] == 0) return 0;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            int v = cell[r][c];
            if ((r + 1 < 4 && cell[r + 1][c] == v) || (c + 1 < 4 && cell[r][c + 1] == v)) return 0;
        }
    return 1;
}

/* Game2048.get_action_mask / _can_change_with_action (src/game2048.py:95-99, :233-237):
 * preview each move on a copy, report whether it changes the board. */
static int action_mask(int cell[4][4]) {
    int mask = 0;
    for (int a = 0; a < 4; ++a) {
        int tmp[4][4], merged[8], nm, ov = 0;
        memcpy(tmp, cell, sizeof(tmp));
        if (board_move(tmp, a, merged, &nm, &ov)) mask |= 1 << a;
    }
    return mask;
}

/* -------------------------------------------------------------- Philox */

/* Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11; constants as in Random123 philox.h). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void stream_words(uint64_t seed, uint64_t gid, uint32_t t, uint32_t domain, uint32_t w[4]) {
    uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), t, domain};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, w);
}

/* --------------------------------------------------------------- spawn */

/* Game2048._spawn (src/game2048.py:108-118) with the device's replayable
 * definition of the two draws: k = mulhi32(wp, n_empty) replaces
 * rng.integers(len(empties)); (wv >= 0xE6666667) replaces rng.random() >= 0.9.
 * Reports the choice so tests can replay it into the reference. */
static void spawn(int cell[4][4], uint32_t wp, uint32_t wv, int* k_out, int* four_out) {
    int n = 0;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) n += cell[r][c] == 0;
    if (k_out) *k_out = -1;
    if (four_out) *four_out = 0;
    if (n == 0) return;
    int k = (int)(((uint64_t)wp * (uint64_t)n) >> 32);
    int four = wv >= 0xE6666667u;
    int seen = 0;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            if (cell[r][c] == 0) {
                if (seen == k) { cell[r][c] = four ? 2 : 1; r = 4; break; }
                ++seen;
            }
    if (k_out) *k_out = k;
    if (four_out) *four_out = four;
}

/* Game2048.reset (src/game2048.py:26-34): empty board, two spawns. */
static void reset_board(int cell[4][4], uint64_t seed, uint64_t gid, uint32_t t, int32_t spawn_log[4]) {
    uint32_t w[4];
    memset(cell, 0, sizeof(int) * 16);
    stream_words(seed, gid, t, B2048_DOM_RESET, w);
    int k, f;
    spawn(cell, w[0], w[1], &k, &f);
    if (spawn_log) { spawn_log[0] = k; spawn_log[1] = f; }
    spawn(cell, w[2], w[3], &k, &f);
    if (spawn_log) { spawn_log[2] = k; spawn_log[3] = f; }
}

int orc_reset_many(uint64_t* board, uint32_t* score, uint32_t* step, uint8_t* max_exp,
                   uint8_t* flags, int32_t* spawn_log /* [n,4] or NULL */, int64_t n,
                   uint64_t seed, uint64_t gid0, uint32_t t) {
    for (int64_t i = 0; i < n; ++i) {
        int cell[4][4];
        reset_board(cell, seed, gid0 + (uint64_t)i, t, spawn_log ? spawn_log + 4 * i : 0);
        board[i] = pack(cell);
        if (score) score[i] = 0;
        if (step) step[i] = 0;
        if (max_exp) max_exp[i] = 2;               /* max_tile_seen = 4 (src/env.py:183) */
        if (flags) flags[i] = (uint8_t)action_mask(cell);
    }
    return 0;
}

/* ------------------------------------------------------------ observation */

/* Game2048Env._preprocess_board (src/env.py:131-150) */
static void encode_obs(int cell[4][4], float* obs, int obs_mode, float log2_scale) {
    if (obs_mode == B2048_OBS_RAW) {
        for (int i = 0; i < 16; ++i) {
            int e = cell[i / 4][i % 4];
            obs[i] = e ? (float)(1u << e) : 0.0f;
        }
    } else if (obs_mode == B2048_OBS_LOG2) {
        for (int i = 0; i < 16; ++i) obs[i] = (float)cell[i / 4][i % 4] * log2_scale;
    } else if (obs_mode == B2048_OBS_ONEHOT) {
        for (int i = 0; i < 16 * 17; ++i) obs[i] = 0.0f;
        for (int i = 0; i < 16; ++i) obs[i * 17 + cell[i / 4][i % 4]] = 1.0f;
    }
}

int orc_encode_obs(const uint64_t* board, float* obs, int32_t obs_mode, float scale, int64_t n) {
    int w = obs_mode == B2048_OBS_ONEHOT ? 272 : 16;
    for (int64_t i = 0; i < n; ++i) {
        int cell[4][4];
        unpack(board[i], cell);
        encode_obs(cell, obs + i * w, obs_mode, scale);
    }
    return 0;
}

/* ------------------------------------------------------------------ step */

/* One environment step: Game2048.step (src/game2048.py:40-70) inside
 * Game2048Env.step (src/env.py:264-302) with _compute_reward (src/env.py:197-261).
 * spawn_log[n,6] (optional): {k, four} of the step spawn (k = -1 if none),
 * then {k1, f1, k2, f2} of the auto-reset spawns ({-1,..} if no reset). */
int orc_step_many(const uint64_t* board_in, uint64_t* board_out,
                  uint32_t* score, uint32_t* step, uint8_t* max_exp,
                  const uint8_t* action, uint8_t* action_out, const uint8_t* flags_in,
                  const b2048_env_cfg* cfg,
                  int32_t* merge_sum, float* reward, double* reward64, uint8_t* flags, float* obs,
                  int32_t* spawn_log, int64_t n, uint64_t seed, uint64_t gid0, uint32_t t) {
    (void)flags_in;
    int obs_w = cfg->obs_mode == B2048_OBS_ONEHOT ? 272 : 16;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t gid = gid0 + (uint64_t)i;
        int cell[4][4];
        unpack(board_in[i], cell);
        uint32_t w[4];
        stream_words(seed, gid, t, B2048_DOM_STEP, w);

        /* action */
        int a;
        if (cfg->action_mode == B2048_ACT_BUFFER) {
            a = action[i] & 3;
        } else if (cfg->action_mode == B2048_ACT_RANDOM_ANY) {
            a = (int)(w[2] >> 30);
        } else if (cfg->action_mode == B2048_ACT_PRIORITY) { /* tools/simple_action_gen.py:16-33: first legal in a fixed order */
            int m = action_mask(cell);
            a = 0;
            for (int k = 3; k >= 0; --k) {
                int c = (cfg->action_priority >> (4 * k)) & 3;
                if (m >> c & 1) a = c;
            }
        } else { /* uniform over legal moves; falls back to 0 if none is legal */
            int m = action_mask(cell), nl = 0, legal[4];
            for (int k = 0; k < 4; ++k)
                if (m >> k & 1) legal[nl++] = k;
            a = nl ? legal[(int)(((uint64_t)w[2] * (uint64_t)nl) >> 32)] : 0;
        }
        if (action_out) action_out[i] = (uint8_t)a;

        /* env.py:267, game2048.py:47 */
        uint32_t stepc = 0;
        if (step) { step[i] += 1; stepc = step[i]; }

        /* game2048.py:49-58 */
        int merged[8], nm = 0, ov = 0;
        int changed = board_move(cell, a, merged, &nm, &ov);
        int32_t msum = 0;
        for (int k = 0; k < nm; ++k) msum += (int32_t)1 << merged[k];
        if (score) score[i] += (uint32_t)msum;
        int sk = -1, sf = 0;
        if (changed) spawn(cell, w[0], w[1], &sk, &sf);
        int done = is_done(cell);

        /* env.py:273 */
        int invalid = !changed && !done;

        /* env.py:197-261, float64, same operation order */
        double r;
        if (!cfg->use_action_mask && invalid) {
            r = cfg->invalid_action_penalty;
        } else {
            if (cfg->reward_mode == B2048_REWARD_SUM) {
                r = (double)msum;
            } else {
                r = 0.0;
                for (int k = 0; k < nm; ++k) r += (double)merged[k];   /* log2(2^e) = e */
            }
            r *= cfg->base_reward_scale;
            if (cfg->empty_tile_reward != 0.0) {
                int ne = 0;
                for (int q = 0; q < 16; ++q) ne += cell[q / 4][q % 4] == 0;
                r += cfg->empty_tile_reward * (double)ne;
            }
            if (cfg->merge_reward != 0.0) r += cfg->merge_reward * (double)nm;
            int mx = 0;
            for (int k = 0; k < nm; ++k)
                if (merged[k] > mx) mx = merged[k];
            int seen = max_exp ? max_exp[i] : 2;
            if (max_exp && mx >= 3 && mx > seen) {
                double bonus = 0.0;
                if (cfg->bonus_mode == B2048_BONUS_RAW) bonus = (double)((uint32_t)1 << mx);
                else if (cfg->bonus_mode == B2048_BONUS_LOG2) bonus = (double)mx;
                max_exp[i] = (uint8_t)mx;
                bonus *= cfg->bonus_scale;
                r += bonus;
            }
            r += cfg->step_reward;
            if (done && cfg->endgame_penalty != 0.0) r += cfg->endgame_penalty;
        }

        /* env.py:279-286 */
        int trunc = step && cfg->max_steps > 0 && stepc >= (uint32_t)cfg->max_steps && !done;

        uint8_t f = (uint8_t)((changed ? B2048_F_CHANGED : 0) | (done ? B2048_F_DONE : 0) |
                              (trunc ? B2048_F_TRUNC : 0) | (ov ? B2048_F_OVERFLOW : 0));
        int32_t rl[4] = {-1, 0, -1, 0};
        if (cfg->auto_reset && (done || trunc)) {
            reset_board(cell, seed, gid, t, rl);
            if (score) score[i] = 0;
            if (step) step[i] = 0;
            if (max_exp) max_exp[i] = 2;
        }
        f |= (uint8_t)action_mask(cell);

        board_out[i] = pack(cell);
        if (merge_sum) merge_sum[i] = msum;
        if (reward) reward[i] = (float)r;
        if (reward64) reward64[i] = r;
        flags[i] = f;
        if (obs && cfg->obs_mode != B2048_OBS_NONE) encode_obs(cell, obs + i * obs_w, cfg->obs_mode, cfg->obs_log2_scale);
        if (spawn_log) {
            int32_t* s = spawn_log + 6 * i;
            s[0] = sk; s[1] = sf; s[2] = rl[0]; s[3] = rl[1]; s[4] = rl[2]; s[5] = rl[3];
        }
    }
    return 0;
}

/* Game2048._move preview (src/game2048.py:158-165) for n boards. */
int orc_move_many(const uint64_t* board_in, uint64_t* board_out, const uint8_t* action,
                  int32_t* merge_sum, uint8_t* merge_info, uint8_t* flags, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        int cell[4][4], merged[8], nm = 0, ov = 0;
        unpack(board_in[i], cell);
        int a = action[i] & 3;
        /* per-line merge bytes, in the line order of board_move */
        uint8_t mi[4] = {0, 0, 0, 0};
        {
            int tmp[4][4];
            memcpy(tmp, cell, sizeof(tmp));
            for (int l = 0; l < 4; ++l) {
                int in[4], out[4], mg[2] = {0, 0}, o2 = 0;
                for (int q = 0; q < 4; ++q) {
                    switch (a) {
                        case 3: in[q] = tmp[l][q]; break;
                        case 1: in[q] = tmp[l][3 - q]; break;
                        case 0: in[q] = tmp[q][l]; break;
                        default: in[q] = tmp[3 - q][l]; break;
                    }
                }
                int k = line_move(in, out, mg, &o2);
                for (int q = 0; q < k; ++q) mi[l] |= (uint8_t)((mg[q] == 16 ? 1 : mg[q]) << (4 * q));
            }
        }
        int changed = board_move(cell, a, merged, &nm, &ov);
        int32_t msum = 0;
        for (int k = 0; k < nm; ++k) msum += (int32_t)1 << merged[k];
        int done = is_done(cell);
        board_out[i] = pack(cell);
        if (merge_sum) merge_sum[i] = msum;
        if (merge_info) memcpy(merge_info + 4 * i, mi, 4);
        if (flags)
            flags[i] = (uint8_t)(action_mask(cell) | (changed ? B2048_F_CHANGED : 0) |
                                 (done ? B2048_F_DONE : 0) | (ov ? B2048_F_OVERFLOW : 0));
    }
    return 0;
}

/* legal mask + done of arbitrary boards (game2048.py:95-99, :172-187) */
int orc_mask_done(const uint64_t* board, uint8_t* mask, uint8_t* done, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        int cell[4][4];
        unpack(board[i], cell);
        if (mask) mask[i] = (uint8_t)action_mask(cell);
        if (done) done[i] = (uint8_t)is_done(cell);
    }
    return 0;
}

/* ------------------------------------------------------------- returns */

/* ReinforceAgent.compute_returns (src/reinforce_agent.py:255-273): float64
 * recurrence G = r + gamma*G, stored as float32.  x,y are [T,B] time-major. */
int orc_reverse_scan(const float* x, float* y, const int32_t* len, double c, int32_t T, int64_t B) {
    for (int64_t b = 0; b < B; ++b) {
        int L = len ? len[b] : T;
        if (L > T) L = T;
        double G = 0.0;
        for (int t = T - 1; t >= 0; --t) {
            if (t >= L) { y[(int64_t)t * B + b] = 0.0f; continue; }
            G = (double)x[(int64_t)t * B + b] + c * G;
            y[(int64_t)t * B + b] = (float)G;
        }
    }
    return 0;
}
