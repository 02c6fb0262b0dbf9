// host_sim.cpp -- TEST INFRASTRUCTURE.  Compiles the device-side source of the CUDA core
// (csrc/a26_core.cuh, policy.cuh, rollout.cuh) for the CPU behind a small intrinsics shim, so the
// kernel logic can be debugged against the oracle in the build container, which has no GPU.  It is
// not part of the product library, is never loaded by the package, and proves nothing about the
// GPU build: the `-m gpu` parity tests do that through libngp.so.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#define __CUDACC__ 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t __brev(uint32_t v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(v);
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline void __syncthreads() {}
static inline unsigned __activemask() { return 1u; }
static inline int __all_sync(unsigned, int v) { return v; }
static inline unsigned __ballot_sync(unsigned, int v) { return v ? 1u : 0u; }
static inline int __syncthreads_or(int v) { return v; }

#include "../../neuro_genetic_pong_self_play_b200/csrc/rollout.cuh"
#include "../../neuro_genetic_pong_self_play_b200/csrc/host_tables.h"

namespace {
// ngp_default_config's button map (csrc/ngp_core.cu)
const uint8_t kDefaultButtonMap[16] = {NGP_BTN_FIRE_P0 + 1, 0, NGP_BTN_SELECT, NGP_BTN_RESET, NGP_BTN_UP_P0 + 2, NGP_BTN_UP_P0 + 3,
                                       NGP_BTN_UP_P0 + 0, NGP_BTN_UP_P0 + 1, 0, 0, NGP_BTN_SELECT, NGP_BTN_RESET, 0, 0, 0, NGP_BTN_FIRE_P0 + 0};
// same palette the product uploads (csrc/ngp_core.cu: ngp_ntsc_palette)
const uint32_t kPalette[128] = {
    0x000000, 0x4a4a4a, 0x6f6f6f, 0x8e8e8e, 0xaaaaaa, 0xc0c0c0, 0xd6d6d6, 0xececec, 0x484800, 0x69690f, 0x86861d, 0xa2a22a,
    0xbbbb35, 0xd2d240, 0xe8e84a, 0xfcfc54, 0x7c2c00, 0x904811, 0xa26221, 0xb47a30, 0xc3903d, 0xd2a44a, 0xdfb755, 0xecc860,
    0x901c00, 0xa33915, 0xb55328, 0xc66c3a, 0xd5824a, 0xe39759, 0xf0aa67, 0xfcbc74, 0x940000, 0xa71a1a, 0xb83232, 0xc84848,
    0xd65c5c, 0xe46f6f, 0xf08080, 0xfc9090, 0x840064, 0x97197a, 0xa8308f, 0xb846a2, 0xc659b3, 0xd46cc3, 0xe07cd2, 0xec8ce0,
    0x500084, 0x68199a, 0x7d30ad, 0x9246c0, 0xa459d0, 0xb56ce0, 0xc57cee, 0xd48cfc, 0x140090, 0x331aa3, 0x4e32b5, 0x6848c6,
    0x7f5cd5, 0x956fe3, 0xa980f0, 0xbc90fc, 0x000094, 0x181aa7, 0x2d32b8, 0x4248c8, 0x545cd6, 0x656fe4, 0x7580f0, 0x8490fc,
    0x001c88, 0x183b9d, 0x2d57b0, 0x4272c2, 0x548ad2, 0x65a0e1, 0x75b5ef, 0x84c8fc, 0x003064, 0x185080, 0x2d6d98, 0x4288b0,
    0x54a0c5, 0x65b7d9, 0x75cceb, 0x84e0fc, 0x004030, 0x18624e, 0x2d8169, 0x429e82, 0x54b899, 0x65d1ae, 0x75e7c2, 0x84fcd4,
    0x004400, 0x1a661a, 0x328432, 0x48a048, 0x5cba5c, 0x6fd26f, 0x80e880, 0x90fc90, 0x143c00, 0x355f18, 0x527e2d, 0x6e9c42,
    0x87b754, 0x9ed065, 0xb4e775, 0xc8fc84, 0x303800, 0x505916, 0x6d762b, 0x88923e, 0xa0ab4f, 0xb7c25f, 0xccd86e, 0xe0ec7c,
    0x482c00, 0x694d14, 0x866a26, 0xa28638, 0xbb9f47, 0xd2b656, 0xe8cc63, 0xfce070,
};

struct Sim {
    a26::Tables T;
    std::vector<uint32_t> needed;
    a26::Snapshot start[2];
    // one stepwise env
    a26::Chip s;
    a26::CpuRegs r;
    uint32_t ram_words[32 * 32];   // lane 0 of a warp-interleaved block
    int players = 2;               // the robot game ('Start') is made with players=1 (main.py:40)
};
}  // namespace

extern "C" {

void *hs_create(const uint8_t *rom)
{
    Sim *sim = new Sim();
    const uint8_t ball[3] = {236, 236, 236}, left[3] = {213, 130, 74}, right[3] = {92, 186, 92};
    ngp_host::build_tables(sim->T, rom, ball, left, right, kPalette);
    sim->needed.resize(a26::TRIGMAX + 1);
    ngp_host::build_paddle_table(sim->needed.data());
    a26::Ram ram{sim->ram_words};
    for (int st = 0; st < 2; ++st) {
        roll::build_start_state(st, sim->s, sim->r, sim->T, ram, sim->needed.data());
        roll::store_snapshot(&sim->start[st], sim->s, sim->r, ram);
    }
    return sim;
}
void hs_destroy(void *h) { delete (Sim *)h; }

void hs_env_reset(void *h, int state)
{
    Sim *sim = (Sim *)h;
    roll::load_snapshot(&sim->start[state], sim->s, sim->r, a26::Ram{sim->ram_words});
    sim->players = state == NGP_STATE_START_1P ? 1 : 2;
}
void hs_env_power_on(void *h)
{
    Sim *sim = (Sim *)h;
    a26::power_on(sim->s, sim->r, sim->T, a26::Ram{sim->ram_words}, sim->needed.data());
    sim->players = 2;
}

// raw console input step (like a26o_run_frame)
int hs_env_run_frame(void *h, int core, int swchb, int fire, int dec, int inc, uint8_t *ram_out, uint8_t *fb, double *loc, uint8_t *valid,
                     uint8_t *regs, uint32_t *digest)
{
    Sim *sim = (Sim *)h;
    a26::Ram ram{sim->ram_words};
    a26::Chip &s = sim->s;
    a26::apply_input(s, sim->needed.data(), (uint32_t)swchb, (uint32_t)fire, (uint32_t)dec, (uint32_t)inc);
    a26::clear_obs(s);
    if (fb) memset(fb, 0, 210 * 160);
    if (core) a26::run_frame_compiled<true, false>(s, sim->r, sim->T, ram, fb);
    else a26::run_frame<true>(s, sim->r, sim->T, ram, fb);
    if (ram_out) for (int i = 0; i < 128; ++i) ram_out[i] = (uint8_t)ram.rd(i);
    if (loc && valid)
        for (int t = 0; t < 3; ++t) {
            valid[t] = s.cnt[t] > 0;
            loc[2 * t] = s.cnt[t] ? (double)s.sy[t] / (double)s.cnt[t] : 0.0;
            loc[2 * t + 1] = s.cnt[t] ? (double)s.sx[t] / (double)s.cnt[t] : 0.0;
        }
    if (regs) {
        const a26::CpuRegs &r = sim->r;
        regs[0] = (uint8_t)r.a; regs[1] = (uint8_t)r.x; regs[2] = (uint8_t)r.y; regs[3] = (uint8_t)r.sp; regs[4] = (uint8_t)a26::pack_p(r);
        regs[5] = (uint8_t)r.pc; regs[6] = (uint8_t)(r.pc >> 8); regs[7] = s.error;
    }
    if (digest) {
        digest[0] = s.cx;
        digest[1] = (uint32_t)s.posp0 | ((uint32_t)s.posp1 << 8) | ((uint32_t)s.posm0 << 16) | ((uint32_t)s.posm1 << 24);
        digest[2] = (uint32_t)s.posbl | ((uint32_t)s.vblank << 8) | ((uint32_t)s.ctrlpf << 16) | ((uint32_t)s.vdelbl << 24);
        digest[3] = (uint32_t)s.charge[0] | ((uint32_t)s.charge[1] << 16);
        digest[4] = (uint32_t)s.charge[2] | ((uint32_t)s.charge[3] << 16);
        digest[5] = (uint32_t)s.grp0_new | ((uint32_t)s.grp1_new << 8) | ((uint32_t)s.enabl_new << 16) | ((uint32_t)s.enabl_old << 24);
        digest[6] = (sim->r.cyc - sim->r.cpu_ls) % a26::LINE_CYCLES;
        digest[7] = s.dump_enabled;
    }
    return s.error;
}

// the fused (no framebuffer) flavour of one frame: what rollout_kernel executes
int hs_env_step_fast(void *h, int core, const uint8_t *action16, uint8_t *ram_out, double *loc, uint8_t *valid)
{
    Sim *sim = (Sim *)h;
    a26::Ram ram{sim->ram_words};
    a26::Chip &s = sim->s;
    const uint32_t in = roll::action_to_input(action16, sim->players, kDefaultButtonMap);
    a26::apply_input(s, sim->needed.data(), in & 0xFF, (in >> 8) & 15, (in >> 12) & 15, (in >> 16) & 15);
    a26::clear_obs(s);
    if (core) a26::run_frame_compiled<false, false>(s, sim->r, sim->T, ram, nullptr);
    else a26::run_frame<false>(s, sim->r, sim->T, ram, nullptr);
    for (int i = 0; i < 128; ++i) ram_out[i] = (uint8_t)ram.rd(i);
    for (int t = 0; t < 3; ++t) {
        valid[t] = s.cnt[t] > 0;
        loc[2 * t] = s.cnt[t] ? (double)s.sy[t] / (double)s.cnt[t] : 0.0;
        loc[2 * t + 1] = s.cnt[t] ? (double)s.sx[t] / (double)s.cnt[t] : 0.0;
    }
    return s.error;
}

int hs_env_step(void *h, int core, const uint8_t *action16, uint8_t *ram_out, uint8_t *fb, double *loc, uint8_t *valid, uint8_t *regs, uint32_t *digest)
{
    const uint32_t in = roll::action_to_input(action16, ((Sim *)h)->players, kDefaultButtonMap);
    return hs_env_run_frame(h, core, in & 0xFF, (int)((in >> 8) & 15), (int)((in >> 12) & 15), (int)((in >> 16) & 15), ram_out, fb, loc, valid, regs, digest);
}

// fused evaluation, lanes executed one after another
void hs_evaluate(void *h, int core, const int32_t *nodes, int n_layers, int bias, int schedule, int games, int max_frames, const float *genomes, int n,
                 const float *hof_genomes, const double *hof_fitness, int n_hof, const int32_t *hof_pick, uint64_t seed,
                 uint64_t generation, double *rewards, int32_t *frames)
{
    Sim *sim = (Sim *)h;
    roll::RolloutParams p;
    memset(&p, 0, sizeof(p));
    p.tables = &sim->T; p.needed = sim->needed.data(); p.start = sim->start;
    p.genomes = genomes; p.hof_genomes = hof_genomes; p.hof_fitness = hof_fitness; p.hof_pick = hof_pick;
    p.n = n; p.n_hof = n_hof; p.games = games; p.schedule = schedule; p.win_score = 3; p.timeout_thresh = 2000; p.max_frames = max_frames;
    p.core = core; p.time_scaler = 100.0; p.paddle_height = 16.0; p.seed = seed; p.generation = generation;
    p.shape.n_layers = n_layers; p.shape.bias = bias;
    int G = 0;
    for (int i = 0; i < n_layers; ++i) p.shape.nodes[i] = nodes[i];
    for (int i = 0; i + 1 < n_layers; ++i) G += (nodes[i] + (bias ? 1 : 0)) * nodes[i + 1];
    p.G = G;
    p.rewards = rewards; p.frames = frames;
    for (int players = 1; players <= 2; ++players)                 // as ngp_evaluate builds it (csrc/ngp_core.cu)
        for (int la = 0; la < 3; ++la)
            for (int ra = 0; ra < 3; ++ra) {
                uint8_t a[16] = {0};
                a[0] = 1; a[15] = 1;
                a[4] = ra == pol::ACT_UP; a[5] = ra == pol::ACT_DOWN;
                a[6] = la == pol::ACT_UP; a[7] = la == pol::ACT_DOWN;
                p.input_table[players - 1][la * 3 + ra] = roll::action_to_input(a, players, kDefaultButtonMap);
            }
    a26::Ram ram{sim->ram_words};
    for (int e = 0; e < n * games; ++e) {
        roll::Episode ep;
        a26::Chip s; a26::CpuRegs r;
        roll::episode_begin(ep, p, e, s, r, ram);
        double reward = 0.0;
        if (core) { while (!roll::episode_frame<1>(ep, p, s, r, sim->T, ram, &reward)) {} }
        else { while (!roll::episode_frame<0>(ep, p, s, r, sim->T, ram, &reward)) {} }
        rewards[e] = reward; frames[e] = ep.frame;
    }
}

// Self-test of the register mirror of the hot latches (csrc/a26_core.cuh: HotLatches, hot_write, hot_display_writes) against
// poke_quick() + the latch effects of tia_apply(), on RANDOM latch states and values -- including states the cartridge never
// visits (delayed players and ball, unlocked missiles).  Returns the number of disagreements (0 = none).
static uint32_t hs_rng(uint64_t &st) { st = st * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(st >> 33); }
// what the write-by-write order does with one write: poke_quick's verdict, and the latch effect either way
static bool hs_reference_write(a26::Chip &c, uint32_t reg, uint32_t v)
{
    if (a26::poke_quick(c, reg, v)) {
        if (reg == 0x0E) c.pf1 = (uint8_t)v;          // equal already: poke_quick leaves these two alone
        if (reg == 0x0F) c.pf2 = (uint8_t)v;
        return true;
    }
    switch (reg) {                                    // tia_apply's latch effects (a26_core.cuh)
    case 0x0D: c.pf0 = (uint8_t)v; break;
    case 0x0E: c.pf1 = (uint8_t)v; break;
    case 0x0F: c.pf2 = (uint8_t)v; break;
    case 0x1B: c.grp0_new = (uint8_t)v; c.grp1_old = c.grp1_new; break;
    case 0x1C: c.grp1_new = (uint8_t)v; c.grp0_old = c.grp0_new; c.enabl_old = c.enabl_new; break;
    case 0x1D: c.enam0 = (uint8_t)v; break;
    case 0x1E: c.enam1 = (uint8_t)v; break;
    default: c.enabl_new = (uint8_t)v; break;
    }
    return false;
}
int hs_hot_latch_selftest(uint64_t seed, int rounds)
{
    int bad = 0;
    uint64_t st = seed * 2 + 1;
    for (int it = 0; it < rounds; ++it) {
        a26::Chip c;
        memset(&c, 0, sizeof(c));
        uint8_t *b = &c.pf0;
        for (int i = 0; i < 16; ++i) {
            const uint32_t r = hs_rng(st);
            // a few distinct values per byte, so that equal/unequal, enabled/disabled and delayed/undelayed all come up often
            const uint8_t pool[6] = {0x00, 0xF0, 0x02, 0x01, 0x03, (uint8_t)(r >> 8)};
            b[i] = pool[r % 6];
        }
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) { const uint32_t r = hs_rng(st); const uint8_t pool[6] = {0x00, 0xF0, 0x02, 0x32, 0xB1, (uint8_t)(r >> 8)}; v[i] = pool[r % 6]; }
        const uint32_t regs[8] = {0x1B, 0x1E, 0x1D, 0x1C, 0x0D, 0x0E, 0x0F, 0x1F};        // the display loop's program order
        // (1) one write at a time, every register
        for (int i = 0; i < 8; ++i) {
            a26::Chip ref = c;
            a26::HotLatches h;
            a26::hot_load(h, c);
            const bool q_ref = hs_reference_write(ref, regs[i], v[i]);
            bool q;
            switch (regs[i]) {
            case 0x0D: q = a26::hot_write<0x0D>(h, v[i]); break;
            case 0x0E: q = a26::hot_write<0x0E>(h, v[i]); break;
            case 0x0F: q = a26::hot_write<0x0F>(h, v[i]); break;
            case 0x1B: q = a26::hot_write<0x1B>(h, v[i]); break;
            case 0x1C: q = a26::hot_write<0x1C>(h, v[i]); break;
            case 0x1D: q = a26::hot_write<0x1D>(h, v[i]); break;
            case 0x1E: q = a26::hot_write<0x1E>(h, v[i]); break;
            default: q = a26::hot_write<0x1F>(h, v[i]); break;
            }
            a26::Chip got = c;
            a26::hot_store(h, got);
            if (q != q_ref || memcmp(&got.pf0, &ref.pf0, 16) != 0) ++bad;
        }
        // (2) the whole iteration at once: all-quick or nothing
        {
            a26::Chip ref = c;
            bool all = true;
            for (int i = 0; i < 8; ++i) all = hs_reference_write(ref, regs[i], v[i]) && all;
            a26::HotLatches h;
            a26::hot_load(h, c);
            const bool ok = a26::hot_display_writes(h, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            a26::Chip got = c;
            a26::hot_store(h, got);
            if (ok != all) ++bad;
            else if (ok && memcmp(&got.pf0, &ref.pf0, 16) != 0) ++bad;
            else if (!ok && memcmp(&got.pf0, &c.pf0, 16) != 0) ++bad;                      // refused: the mirror must be untouched
        }
    }
    return bad;
}

}  // extern "C"
