/* episode_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's per-frame control path around the emulator:
 *   utils.find_stuff / get_rect_quickly      /root/reference/utils.py:14-19, 60-68
 *   utils.inference                          /root/reference/utils.py:139-153
 *   numpy_nn.NeuralNetwork.run / sigmoid     /root/reference/numpy_nn.py:120-137, 22-23
 *   dumb_ais.HardcodedAi / ScoreHardcodedAi  /root/reference/dumb_ais.py:1-25
 *   utils.keep_within_game_bounds_please     /root/reference/utils.py:71-77
 *   main.get_actions                         /root/reference/main.py:138-154
 *   main.calculate_timeout_and_frames        /root/reference/main.py:128-135
 *   main.perform_episode                     /root/reference/main.py:69-112
 *   utils.calculate_reward                   /root/reference/utils.py:104-109
 *   main.evaluate                            /root/reference/main.py:28-66
 * The numpy pieces are pinned against the imported reference by tools/make_golden.py
 * (fixtures in tests/golden/).  The emulator underneath is a2600_oracle.c (parity
 * unpinned against Stella; see its header).
 */
#ifndef EPISODE_ORACLE_H
#define EPISODE_ORACLE_H
#include <stdint.h>
#include "a2600_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* config.py mirror */
#define EO_GAME_TOP 34
#define EO_GAME_BOTTOM 194
#define EO_GAME_WIDTH 160
#define EO_PLAYABLE_HEIGHT 160
#define EO_PADDLE_HEIGHT 16.0
#define EO_TIMEOUT_THRESH 2000
#define EO_WIN_SCORE 3
#define EO_TIME_SCALER 100.0
#define EO_GAMES_TO_PLAY 6

enum { EO_ACT_NONE = 0, EO_ACT_UP = 1, EO_ACT_DOWN = 2 };              /* [0,0] [1,0] [0,1] */
enum { EO_POLICY_HARDCODED = 0, EO_POLICY_SCORE_HARDCODED = 1, EO_POLICY_MLP = 2 };
enum { EO_STATE_START_1P = 0, EO_STATE_START_2P = 1 };

typedef struct {
    int n_layers;           /* number of node layers, e.g. 3 for [6,2,2] */
    int nodes[8];
    int bias;               /* config.BIAS */
} eo_shape;

typedef struct {
    int kind;
    const float *genome;    /* EO_POLICY_MLP only */
} eo_policy;

typedef struct {
    double loc[3][2];       /* ball, left, right: (row, col) in cropped coordinates */
    uint8_t valid[3];       /* 0 == the reference's None */
} eo_obs;

typedef struct {
    int frames;             /* env.step calls made */
    int score1, score2;
    double total_frames;    /* the reference's total_frames accumulator */
    double reward;
} eo_episode_result;

int eo_gene_size(const eo_shape *);                                         /* utils.py:128-136 */
double eo_det_exp(double x);                                                /* deterministic exp shared with the CUDA core */
void eo_mlp_forward(const eo_shape *, const float *genome, const double *x, double *out, int *action);
void eo_find_stuff(const uint8_t *rgb /*[210][160][3]*/, eo_obs *out);
int eo_clamp(int valid, double paddle_row, int action);
double eo_reward(double mult, double total_frames, int my_score, int enemy_score);
int eo_bot_act(int kind, const double x[6], int score1, int score2);         /* dumb_ais.py:1-25 */
void eo_inference_vector(const double ball[2], const double last[2], double me_row, double enemy_row, double x[6]);   /* utils.py:139-153 */
uint32_t eo_philox_bit(uint64_t seed, uint64_t generation, uint32_t env_id, uint32_t frame, uint32_t stream);
void eo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* button map codes (one per gym-retro button; the product's NGP_BTN_* in include/ngp.h) */
enum { EO_BTN_NONE = 0, EO_BTN_FIRE_P0 = 1, EO_BTN_UP_P0 = 5, EO_BTN_SELECT = 13, EO_BTN_RESET = 14 };
void eo_set_button_map(const uint8_t map[16]);
void eo_get_button_map(uint8_t map[16]);
/* retro action[16] -> console input for an env made with `players` players (gym-retro button layout) */
void eo_action_to_input(const uint8_t action[16], int players, a26o_input *in);
/* power-on + scripted switches up to the reference's 'Start' / 'Start.2P' save states,
 * including gym-retro reset()'s one idle frame */
void eo_reset_to_state(a26o *env, int state_id);

/* One perform_episode.  trace (optional) receives per frame 144 bytes:
 * RAM[128] | left_act right_act score1 score2 | valid[3] pad | timeout(u32) frame(u32) */
void eo_episode(a26o *env, const eo_shape *shape, eo_policy left, eo_policy right, double mult,
                uint64_t seed, uint64_t generation, uint32_t env_id, int players, int max_frames, eo_episode_result *res,
                uint8_t *trace, int trace_cap);

/* main.evaluate for one genome.  hof_pick[3]: HoF member index for games 3..5 (ignored when
 * n_hof == 0).  Returns the mean reward; rewards[6] and frames[6] are filled. */
double eo_evaluate(const uint8_t rom[2048], const eo_shape *shape, const float *genome,
                   const float *hof_genomes, const double *hof_fitness, int n_hof, const int hof_pick[3],
                   uint64_t seed, uint64_t generation, uint32_t genome_id, double rewards[EO_GAMES_TO_PLAY], int frames[EO_GAMES_TO_PLAY]);

/* round-robin self-play game: right = genome a, left = genome b, 2-player start state */
void eo_selfplay_game(const uint8_t rom[2048], const eo_shape *shape, const float *right, const float *left,
                      uint64_t seed, uint64_t generation, uint32_t env_id, eo_episode_result *res);

#ifdef __cplusplus
}
#endif
#endif
