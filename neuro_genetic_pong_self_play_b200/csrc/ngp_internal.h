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
