/* a2600_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Clear, single-environment, per-colour-clock restatement of the Atari 2600 machine
 * (6507 + TIA + RIOT + paddle controllers) that the reference drives through
 * gym-retro's Stella core at /root/reference/main.py:77 (env.step), :56/:108 (env.reset)
 * and :21/:40/:51 (retro.make).  gym-retro / Stella are third-party dependencies that are
 * NOT vendored in /root/reference (requirements.txt:3, unpinned) and cannot be installed
 * here, so this file follows public 2600 hardware behaviour plus the Stella-3.x
 * conventions listed in DESIGN.md ("Emulator spec").  PARITY UNPINNED against Stella:
 * the only in-tree pins are obs.npy's frame geometry/colours and config.py:3-6.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use anything in oracle/.
 */
#ifndef A2600_ORACLE_H
#define A2600_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A26O_FB_ROWS 210
#define A26O_FB_COLS 160
#define A26O_YSTART 34           /* first framebuffer row = scanline 34 after VSYNC end */
#define A26O_TRIGMAX 4096        /* paddle charge range (Stella Paddles TRIGMAX) */

typedef struct {
    uint8_t swchb;      /* console switches as read at SWCHB (active low reset/select) */
    uint8_t fire;       /* bit i set = paddle i fire button held (i=0..3) */
    uint8_t dec;        /* bit i set = paddle i "decrease charge" key held (screen-up)  */
    uint8_t inc;        /* bit i set = paddle i "increase charge" key held (screen-down) */
} a26o_input;

typedef struct a26o a26o;

/* needed[c] = CPU cycles after the dump is released until INPTx bit7 goes high for paddle
 * charge c (0..4096); built by a26o_build_paddle_table. */
void a26o_build_paddle_table(uint32_t needed[A26O_TRIGMAX + 1]);

a26o *a26o_new(const uint8_t rom[2048]);
void a26o_free(a26o *);
void a26o_power_on(a26o *);
/* Run one frame (until the instruction that turns VSYNC off).  fb may be NULL; otherwise
 * receives 210*160 colour-register values (COLUxx value & 0xFE; 0 = blank/black).
 * Returns 0, or a negative error (illegal opcode, decimal mode, unsupported TIA use). */
int a26o_run_frame(a26o *, const a26o_input *in, uint8_t *fb);
const uint8_t *a26o_ram(const a26o *);              /* 128 bytes */
void a26o_cpu_regs(const a26o *, uint8_t out[8]);   /* A X Y SP P PCL PCH 0 */
uint64_t a26o_cycles(const a26o *);
uint64_t a26o_instructions(const a26o *);
int a26o_state_size(void);
void a26o_save(const a26o *, void *dst);
void a26o_load(a26o *, const void *src);
/* a frame-level digest of TIA-visible state for parity checks (collision latches,
 * object positions, paddle charges) */
void a26o_tia_digest(const a26o *, uint32_t out[8]);

/* Stella NTSC palette (128 entries, index = colour value >> 1), 0x00RRGGBB. */
extern const uint32_t a26o_ntsc_palette[128];
/* palette-index frame -> RGB uint8[210][160][3] (the obs.npy layout) */
void a26o_fb_to_rgb(const uint8_t *fb, uint8_t *rgb);

#ifdef __cplusplus
}
#endif
#endif
