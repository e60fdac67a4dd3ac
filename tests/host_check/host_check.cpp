// host_check.cpp — TEST INFRASTRUCTURE.  Compiles the product's device headers (b2048_device.cuh,
// b2048_step.cuh are host+device) with g++ so the bit tricks / step body can be checked against the
// CPU oracle in the GPU-less container.  Not part of the product; never loaded by the package.
#include <cstdint>
#include <vector>
#include "../../rl-2048-with-reinforce-and-actor-critic_b200/csrc/b2048_step_fast.cuh"

static std::vector<uint16_t> g_left;
static std::vector<uint8_t> g_merge;

static void ensure_lut() {
    if (!g_left.empty()) return;
    g_left.resize(65536);
    g_merge.resize(65536);
    for (uint32_t r = 0; r < 65536; ++r) {
        uint32_t o, m;
        b2::row_move_left(r, o, m);
        g_left[r] = (uint16_t)o;
        g_merge[r] = (uint8_t)m;
    }
}

extern "C" {

void hc_get_lut(uint16_t* left, uint8_t* merge) {
    ensure_lut();
    for (int i = 0; i < 65536; ++i) { left[i] = g_left[i]; merge[i] = g_merge[i]; }
}

void hc_move_many(const uint64_t* in, uint64_t* out, const uint8_t* action, int32_t* merge_sum, uint8_t* flags, int64_t n) {
    ensure_lut();
    for (int64_t i = 0; i < n; ++i) {
        b2::Board b = b2::make_board(in[i]);
        b2::MoveResult mv = b2::move_board(b, action[i] & 3u, g_left.data(), g_merge.data());
        b2::MergeStats ms = b2::merge_stats(mv.merge, true, true);
        out[i] = b2::to_u64(mv.board);
        merge_sum[i] = (int32_t)ms.sum;
        uint32_t mask = b2::legal_mask(mv.board);
        bool done = mask == 0 && out[i] != 0;
        flags[i] = (uint8_t)(mask | (out[i] != in[i] ? B2048_F_CHANGED : 0) | (done ? B2048_F_DONE : 0) |
                             (ms.overflow ? B2048_F_OVERFLOW : 0));
    }
}

void hc_mask(const uint64_t* in, uint8_t* mask, int64_t n) {
    for (int64_t i = 0; i < n; ++i) mask[i] = (uint8_t)b2::legal_mask(b2::make_board(in[i]));
}

void hc_reset_many(uint64_t* board, uint8_t* flags, int64_t n, uint64_t seed, uint64_t gid0, uint32_t t) {
    for (int64_t i = 0; i < n; ++i) {
        b2::Board b = b2::reset_board(seed, gid0 + (uint64_t)i, t);
        board[i] = b2::to_u64(b);
        flags[i] = (uint8_t)b2::legal_mask(b);
    }
}

void hc_step_many(const uint64_t* board_in, uint64_t* board_out, uint32_t* score, uint32_t* step, uint8_t* max_exp,
                  const uint8_t* action, uint8_t* action_out, const uint8_t* flags_in, const b2048_env_cfg* cfg,
                  int32_t* merge_sum, float* reward, double* reward64, uint8_t* flags, int64_t n, uint64_t seed,
                  uint64_t gid0, uint32_t t) {
    ensure_lut();
    b2::StepOpts opt{step != nullptr, max_exp != nullptr, true};
    for (int64_t i = 0; i < n; ++i) {
        b2::StepIO io;
        io.board = b2::make_board(board_in[i]);
        io.score = score ? score[i] : 0;
        io.step = step ? step[i] : 0;
        io.max_exp = max_exp ? max_exp[i] : 2;
        io.action = action ? action[i] : 0;
        io.have_mask_in = flags_in != nullptr;
        io.mask_in = flags_in ? flags_in[i] : 0;
        io.replay = 0;
        b2::step_one(io, *cfg, opt, seed, gid0 + (uint64_t)i, t, g_left.data(), g_merge.data());
        board_out[i] = b2::to_u64(io.board);
        if (score) score[i] = io.score;
        if (step) step[i] = io.step;
        if (max_exp) max_exp[i] = (uint8_t)io.max_exp;
        if (action_out) action_out[i] = (uint8_t)io.action_played;
        if (merge_sum) merge_sum[i] = io.merge_sum;
        if (reward) reward[i] = (float)io.reward;
        if (reward64) reward64[i] = io.reward;
        flags[i] = (uint8_t)io.flags;
    }
}
}  // extern "C"

static std::vector<b2::AggEntry> g_agg;
static std::vector<b2::SelEntry> g_sel;
static std::vector<uint8_t> g_act;

static void ensure_small() {
    if (!g_agg.empty()) return;
    g_agg.resize(256); g_sel.resize(4); g_act.resize(64);
    for (uint32_t b = 0; b < 256; ++b) b2::small_table_entry_agg(b, g_agg[b]);
    for (uint32_t a = 0; a < 4; ++a) b2::small_table_entry_sel(a, g_sel[a]);
    for (uint32_t m = 0; m < 16; ++m)
        for (uint32_t j = 0; j < 4; ++j) g_act[m * 4 + j] = (uint8_t)b2::small_table_entry_act(m, j);
}

template <int kAct, bool kTrack>
static void run_fast(const uint64_t* board_in, uint64_t* board_out, uint32_t* score, uint32_t* step, uint8_t* max_exp,
                     const uint8_t* action, uint8_t* action_out, const uint8_t* flags_in, const b2048_env_cfg* cfg,
                     int32_t* merge_sum, float* reward, uint8_t* flags, int64_t n, uint64_t seed, uint64_t gid0, uint32_t t) {
    b2::FastTables T{g_left.data(), g_merge.data(), g_agg.data(), g_sel.data(), g_act.data()};
    const b2::PhiloxKeys keys = b2::make_keys(seed);
    for (int64_t i = 0; i < n; ++i) {
        b2::FastIO io;
        io.lo = (uint32_t)board_in[i]; io.hi = (uint32_t)(board_in[i] >> 32);
        io.score = kTrack ? score[i] : 0; io.step = kTrack ? step[i] : 0; io.max_exp = kTrack ? max_exp[i] : 2;
        io.action = action ? action[i] : 0;
        io.mask_in = flags_in ? flags_in[i] : b2::legal_mask(b2::Board{io.lo, io.hi});
        b2::step_fast<kAct, kTrack>(io, *cfg, keys, seed, gid0 + (uint64_t)i, t, T);
        board_out[i] = (uint64_t)io.lo | ((uint64_t)io.hi << 32);
        if (kTrack) { score[i] = io.score; step[i] = io.step; max_exp[i] = (uint8_t)io.max_exp; }
        if (action_out) action_out[i] = (uint8_t)io.action;
        if (merge_sum) merge_sum[i] = io.merge_sum;
        if (reward) reward[i] = io.reward;
        flags[i] = (uint8_t)io.flags;
    }
}

extern "C" void hc_step_fast_many(const uint64_t* board_in, uint64_t* board_out, uint32_t* score, uint32_t* step, uint8_t* max_exp,
                       const uint8_t* action, uint8_t* action_out, const uint8_t* flags_in, const b2048_env_cfg* cfg,
                       int32_t* merge_sum, float* reward, uint8_t* flags, int64_t n, uint64_t seed, uint64_t gid0,
                       uint32_t t) {
    ensure_lut(); ensure_small();
    bool track = score && step && max_exp;
#define RUN(A) (track ? run_fast<A, true>(board_in, board_out, score, step, max_exp, action, action_out, flags_in, cfg, merge_sum, reward, flags, n, seed, gid0, t) \
                      : run_fast<A, false>(board_in, board_out, score, step, max_exp, action, action_out, flags_in, cfg, merge_sum, reward, flags, n, seed, gid0, t))
    if (cfg->action_mode == B2048_ACT_BUFFER) RUN(B2048_ACT_BUFFER);
    else if (cfg->action_mode == B2048_ACT_RANDOM_ANY) RUN(B2048_ACT_RANDOM_ANY);
    else if (cfg->action_mode == B2048_ACT_PRIORITY) RUN(B2048_ACT_PRIORITY);
    else RUN(B2048_ACT_RANDOM_LEGAL);
#undef RUN
}
