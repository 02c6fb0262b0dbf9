// ngp_internal.h -- handle layout and helpers shared by the translation units of libngp.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ngp.h"
#include "a26_core.cuh"
#include "policy.cuh"

struct ngp_handle {
    ngp_config cfg;
    int device;
    int sm_count;
    int gene_size;
    pol::Shape shape;
    int rom_translated;           // the ROM is the cartridge the statically translated core was generated from
    // device tables
    a26::Tables *d_tables;        // rom + decode + colour-match weights
    uint32_t *d_needed;           // paddle charge -> cycles table [4097]
    uint32_t *d_palette;          // NTSC palette [128]
    a26::Snapshot *d_start;       // [2]: 'Start' (1P) and 'Start.2P'
    // stepwise (explicit action) environments
    a26::Snapshot *d_envs;
    uint8_t *d_fb;                // palette-index frames [n_envs][210][160]
    int n_envs, cap_envs;
    int env_players;              // gym-retro `players` of the state the stepwise environments were reset to (1 for 'Start')
    // fused evaluation scratch
    double *d_rewards;            // [cap_eval]
    int32_t *d_frames;            // [cap_eval]
    unsigned long long *d_counters;   // [0] next env, [1] frames, [2] emulator errors
    int cap_eval;
    // host staging (pinned) for the *_host entry points
    float *h_genomes; double *h_fitness; float *d_genomes_stage; double *d_fitness_stage;
    size_t stage_genomes, stage_fitness;
    float *d_hof_stage; double *d_hof_fit_stage; size_t stage_hof;
    unsigned long long *h_counters;
    uint64_t launches;
    // per-handle scratch of the operator entry points (nothing here is shared between handles)
    float *mlp_a, *mlp_b; double *mlp_z; size_t mlp_cap_ab, mlp_cap_z;      // ngp_mlp_forward activations
    int32_t *d_parent; size_t cap_parent;                                   // ngp_ga_step selection winners
    void *rank_keys; int32_t *rank_idx; size_t cap_rank;                    // fitness ranking of the order-statistics tournament
    uint64_t *hof_hash_old, *hof_hash_new; int32_t *hof_order; float *hof_tmp_genomes; double *hof_tmp_fitness;   // ngp_hof_update
    size_t hof_cap_hash_old, hof_cap_hash_new, hof_cap_order, hof_cap_tmp_genomes, hof_cap_tmp_fitness;
    // ngp_evaluate for nets wider than the fused rollout (ngp_stepwise.cu): per-environment state, MLP input rows, actions,
    // gathered hall-of-fame opponents
    void *step_envs; float *step_x; uint8_t *step_act; float *step_opp;
    size_t step_cap_envs, step_cap_x, step_cap_act, step_cap_opp;
    // ngp_mlp_prepare: packed wide layers of the prepared genome set (ngp_mlp_tmem.cu)
    float *prep_packed; size_t prep_cap, prep_per_genome; const float *prep_src; int prep_n; int tmem_attr_set;
    int fs_per_sm;                                                          // resident find_stuff CTAs per SM
    int tf32_attr_set;                                                      // dynamic shared memory opt-in done
    // ngp_set_option (tuning experiments; 0 = automatic)
    int opt_rollout_block, opt_rollout_nosync, opt_rollout_lean, opt_rollout_blocks_per_sm, opt_mlp_no_tf32, opt_select_os_min_t;
    // profiling of the rollout kernel
    int profile_on;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *prof_events;
    int prof_used;
};

void ngp_set_error(const char *fmt, ...);
#define NGP_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            ngp_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return NGP_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)
#define NGP_REQUIRE(cond, msg)                                      \
    do {                                                            \
        if (!(cond)) { ngp_set_error("%s", msg); return NGP_ERR_INVALID; } \
    } while (0)

extern const uint32_t ngp_ntsc_palette[128];
